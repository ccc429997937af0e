#!/bin/bash
# round-2 experiment batch Q: searches that look `slack` beyond the best distance (better bounds for the correspondence
# cache), sqrt-free cache test
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2q_tests.log
run() { # suffix lanes contexts tag
  PTK_LIB_SUFFIX=$1 timeout 300 python bench.py --lanes $2 --contexts $3 --no-side-runs --no-cpu-baseline --no-e2e \
     > $O/r2q_v$1_l$2c$3$4.json 2> $O/r2q_v$1_l$2c$3$4.err; echo "v$1 l$2 c$3 $4 rc=$?"
}
run "" 48 1
run _sl01 48 1
run _sl02 48 1
run _sl035 48 1
run "" 64 8
run _sl01 64 8
run _sl02 64 8
run _sl035 64 8
