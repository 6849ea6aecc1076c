#!/bin/bash
# usage: scripts/ship.sh "<commit message>" <gpu-timeout> <remote command>
# commit the dev worktree, fast-forward main, rebuild main's library, run the command on a GPU box (with retries)
set -e
MSG="$1"; T="$2"; shift 2
git -C /root/repo/.wt/dev add -A
git -C /root/repo/.wt/dev commit -q -m "$MSG" || true
git -C /root/repo merge -q dev
(cd /root/repo && python -c "import __graft_entry__ as g; g.build()")
git -C /root/repo log --oneline | head -1
cd /root/repo && scripts/gpurun_retry.sh "$T" "$@"
