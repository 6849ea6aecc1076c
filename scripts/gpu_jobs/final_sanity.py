"""Torch-free sanity check of the shipped libosb200.so on a GPU box (a few seconds: ctypes + libcudart + NumPy only).
device info, SM clock kernel, fp32 <-> split-bf16 round trip, a 1x1x1 fp32 convolution against NumPy."""
import ctypes
import sys
import time

import numpy as np

t0 = time.time()
L = ctypes.CDLL('openscene_b200/libosb200.so')      # pulls libcudart.so.12 in through its own dependency
rt = ctypes.CDLL('libcudart.so.12')
L.osb_last_error.restype = ctypes.c_char_p
P, I32, I64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64


def chk(rc, what):
    if rc:
        print('FAIL', what, rc, L.osb_last_error())
        sys.exit(1)


def dmalloc(nbytes):
    p = P()
    chk(rt.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(nbytes)), 'cudaMalloc')
    return p


def h2d(arr):
    p = dmalloc(arr.nbytes)
    chk(rt.cudaMemcpy(p, arr.ctypes.data_as(P), ctypes.c_size_t(arr.nbytes), 1), 'h2d')
    return p


def d2h(p, shape, dtype):
    out = np.empty(shape, dtype)
    chk(rt.cudaMemcpy(out.ctypes.data_as(P), p, ctypes.c_size_t(out.nbytes), 2), 'd2h')
    return out


sm, maj, mnr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
chk(L.osb_device_info(ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mnr)), 'osb_device_info')
print(f'device: {sm.value} SMs, sm_{maj.value}{mnr.value}  ({time.time() - t0:.2f} s)')
mhz = dmalloc(4)
L.osb_measure_sm_mhz.argtypes = [P, P]
chk(L.osb_measure_sm_mhz(mhz, None), 'osb_measure_sm_mhz')
chk(rt.cudaDeviceSynchronize(), 'sync')
print('sm clock MHz', float(d2h(mhz, (1,), np.float32)[0]))

rng = np.random.RandomState(0)
x = rng.randn(300, 64).astype(np.float32)
dx, ds, dy = h2d(x), dmalloc(x.nbytes), dmalloc(x.nbytes)
L.osb_f32_to_split.argtypes = [P, I64, I32, P, P]
L.osb_split_to_f32.argtypes = [P, I64, I32, P, P]
chk(L.osb_f32_to_split(dx, 300, 64, ds, None), 'osb_f32_to_split')
chk(L.osb_split_to_f32(ds, 300, 64, dy, None), 'osb_split_to_f32')
chk(rt.cudaDeviceSynchronize(), 'sync')
y = d2h(dy, x.shape, np.float32)
err = float(np.max(np.abs(y - x) / np.abs(x)))
print('split round trip max rel err', err)
assert err <= 2.0 ** -16

w = rng.randn(1, 64, 48).astype(np.float32)
dw, do = h2d(w), dmalloc(300 * 48 * 4)
L.osb_conv_fwd_f32.argtypes = [P, I64, P, I64, I32, P, I32, I32, I32, P, P]
chk(L.osb_conv_fwd_f32(dx, 64, None, 300, 1, dw, 64, 48, 0, do, None), 'osb_conv_fwd_f32')
chk(rt.cudaDeviceSynchronize(), 'sync')
o = d2h(do, (300, 48), np.float32)
ref = x.astype(np.float64) @ w[0].astype(np.float64)
e2 = float(np.max(np.abs(o - ref)) / np.max(np.abs(ref)))
print('1x1x1 conv vs numpy', e2)
assert e2 < 1e-5
print(f'SANITY OK in {time.time() - t0:.2f} s')
