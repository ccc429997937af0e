#!/bin/bash
# round-2 experiment batch N: column-per-thread k_scan_insert_range + candidate-only slot1 + compacted winners in k_compact1
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2n_tests.log
run() { # suffix lanes contexts
  PTK_LIB_SUFFIX=$1 timeout 300 python bench.py --lanes $2 --contexts $3 --no-side-runs --no-cpu-baseline --no-e2e \
     > $O/r2n_v$1_l$2c$3.json 2> $O/r2n_v$1_l$2c$3.err; echo "v$1 l$2 c$3 rc=$?"
}
run "" 64 8
run "" 48 1
run _si4 64 8
run _r2 64 8
run _r8 64 8
