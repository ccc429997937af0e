#!/bin/bash
# round-2 batch Y: ICP kernels on a high-priority side stream (PTK_ICP_PRIORITY=1) against the default
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
run() { # prio tag
  PTK_ICP_PRIORITY=$1 timeout 300 python bench.py --no-side-runs --no-cpu-baseline --no-e2e > $O/r2y_p$1$2.json 2> $O/r2y_p$1$2.err; echo "p$1 $2 rc=$?"
}
run 0 a
run 1 a
run 0 b
run 1 b
PTK_ICP_PRIORITY=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch or fleet or golden or wrapper" > $O/r2y_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/r2y_tests.log
