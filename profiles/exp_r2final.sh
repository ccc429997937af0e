#!/bin/bash
# round-2 final check of HEAD: GPU tests, smoke(), default bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2f_tests.log
timeout 300 python __graft_entry__.py --smoke > $O/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2f_smoke.log
timeout 600 python bench.py > $O/r2f_bench.json 2> $O/r2f_bench.err; echo "bench rc=$?"
