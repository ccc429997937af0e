#!/bin/bash
# round-2 experiment batch M: validation of HEAD + ICP block-shape variants (more lanes in flight with smaller blocks)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2m_tests.log 2>&1; echo "tests rc=$?"
timeout 600 python bench.py > $O/r2m_bench.json 2> $O/r2m_bench.err; echo "bench rc=$?"
timeout 300 taskset -c 0-3 python bench.py --no-side-runs --no-cpu-baseline > $O/r2m_bench_4cores.json 2> $O/r2m_bench_4cores.err; echo "bench4 rc=$?"
run() { # suffix lanes contexts blocks
  PTK_LIB_SUFFIX=$1 timeout 300 python bench.py --lanes $2 --contexts $3 --icp-blocks $4 --no-side-runs --no-cpu-baseline --no-e2e \
     > $O/r2m_v$1_l$2c$3b$4.json 2> $O/r2m_v$1_l$2c$3b$4.err; echo "v$1 l$2 c$3 b$4 rc=$?"
}
run "" 96 12 6
run "" 128 16 6
run _t256c416 64 8 5
run _t256c416 96 12 5
run _t256c416 128 16 5
run _t256c352 64 8 6
run _t256c352 96 12 6
run _t256c352 128 16 6
run _t256c352 96 12 7
run _t320c320 64 8 7
run _t320c320 96 12 7
