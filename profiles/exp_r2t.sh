#!/bin/bash
# round-2 batch T: decode kernel grid shapes; the new config-3 long-run test; ingest tests
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 300 python profiles/run_ingest.py "" 0 1 2 4 > $O/r2t_ingest.log 2>&1; echo "ingest rc=$?"; cat $O/r2t_ingest.log | tail -6
timeout 600 python -m pytest tests/test_ingest.py "tests/test_gpu_parity.py::test_os2_long_run_with_prune_equals_the_cpu_port" -m gpu -x -q > $O/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2t_tests.log
