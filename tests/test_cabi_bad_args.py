"""Error behaviour at the C ABI ("returns int status with a thread-local error string; no exception -- and no
signal -- crosses the ABI").  Every entry point is called with all-NULL / all-zero and with negative arguments, in a child
process so that a crash would show up as a failed test instead of taking pytest down.  No GPU is needed: argument
validation happens on the host before the first CUDA call."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r'''
import ctypes, json, sys
sys.path.insert(0, sys.argv[1])
from openscene_b200 import _cabi as C
from ctypes import c_char_p, c_double, c_void_p
L = C.lib()
out = {}
for name, (res, argt) in sorted(C.SIGNATURES.items()):
    for mode in ("null", "neg"):
        args = []
        for t in argt:
            if t is c_void_p or t is c_char_p or (isinstance(t, type) and issubclass(t, ctypes._Pointer)):
                args.append(None)
            elif t is c_double:
                args.append(0.0)
            else:
                args.append(0 if mode == "null" else -1)
        r = getattr(L, name)(*args)
        err = L.osb_last_error()
        out[name + ":" + mode] = [int(r) if isinstance(r, int) else None, (err or b"").decode(errors="replace")]
print("RESULT " + json.dumps(out))
'''

# not compute entry points: constants, counters, size queries (any value is legal; they must only survive the call)
_QUERIES = {'osb_version', 'osb_last_error', 'osb_launch_count', 'osb_conv_chain_grid', 'osb_conv_desc_bytes', 'osb_device_info'}


def test_degenerate_arguments_fail_with_a_message_and_never_crash():
    p = subprocess.run([sys.executable, '-c', _CHILD, ROOT], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, f"the library crashed on degenerate arguments (exit {p.returncode}):\n{p.stderr[-2000:]}"
    line = [l for l in p.stdout.splitlines() if l.startswith('RESULT ')][-1]
    res = json.loads(line[len('RESULT '):])
    from openscene_b200 import _cabi as C
    assert len(res) == 2 * len(C.SIGNATURES)
    for key, (rc, err) in res.items():
        name, mode = key.split(':')
        if name in _QUERIES or name.endswith('_bytes'):
            continue
        if name == 'osb_gather_rows_f32' and mode == 'null':
            assert rc == 0                                   # zero rows: a documented no-op
            continue
        assert rc != 0, f"{key}: accepted degenerate arguments"
        assert err.strip(), f"{key}: failed without a message"
    # size queries of rejected shapes reserve nothing
    for name in ('osb_conv_chain_workspace_bytes', 'osb_conv_tc_workspace_bytes', 'osb_conv_wgrad_tc_workspace_bytes',
                 'osb_conv_packed_weight_bytes', 'osb_conv_weight_tiles_bytes', 'osb_occgrid_bytes'):
        assert res[name + ':null'][0] == 0 and res[name + ':neg'][0] == 0, name
