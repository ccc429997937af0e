#!/bin/bash
# round-2 batch X: final state - GPU tests, smoke(), default bench (timed), reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2x_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2x_tests.log
timeout 300 python __graft_entry__.py --smoke > $O/r2x_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2x_smoke.log
s=$(date +%s); timeout 600 python bench.py > $O/r2x_bench.json 2> $O/r2x_bench.err; echo "bench rc=$? in $(( $(date +%s) - s )) s"
s=$(date +%s); timeout 600 python bench.py --impl reference > $O/r2x_ref.json 2> $O/r2x_ref.err; echo "ref rc=$? in $(( $(date +%s) - s )) s"
