#!/bin/bash
# round-2 experiment batch P: default build (full-warp searches, new streaming kernels): tests, full bench with side runs
# (first run of the ingest side run); solve-phase clocks (_sc)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2p_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2p_tests.log
timeout 600 python bench.py > $O/r2p_bench.json 2> $O/r2p_bench.err; echo "bench rc=$?"
PTK_LIB_SUFFIX=_sc timeout 300 python bench.py --lanes 48 --contexts 1 --no-side-runs --no-cpu-baseline --no-e2e > $O/r2p_sc_l48c1.json 2> $O/r2p_sc_l48c1.err; echo "sc rc=$?"
PTK_LIB_SUFFIX=_sc timeout 300 python bench.py --lanes 64 --contexts 8 --no-side-runs --no-cpu-baseline --no-e2e > $O/r2p_sc_l64c8.json 2> $O/r2p_sc_l64c8.err; echo "sc rc=$?"
