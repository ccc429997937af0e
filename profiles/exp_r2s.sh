#!/bin/bash
# round-2 batch S: validation + evidence for the final kernels: GPU tests, compute-sanitizer, default bench + reference
# arm, ncu launch list, ncu --set full capture of one step
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2s_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2s_tests.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_ingest.py tests/test_golden.py \
  "tests/test_gpu_parity.py::test_register_scan_range_image_path" "tests/test_gpu_parity.py::test_register_scan_batch_matches_single" \
  "tests/test_gpu_parity.py::test_empty_and_degenerate_scans" "tests/test_gpu_parity.py::test_wrapper_with_extrinsics_and_offsets" \
  -m gpu -x -q > $O/r2s_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/r2s_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_ingest.py \
  "tests/test_gpu_parity.py::test_register_scan_range_image_path" -m gpu -x -q > $O/r2s_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 $O/r2s_racecheck.log
timeout 600 python bench.py > $O/r2s_bench.json 2> $O/r2s_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/r2s_ref.json 2> $O/r2s_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:k_ -c 1500 --csv --log-file $O/r2s_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-side-runs --no-cpu-baseline --no-e2e > $O/r2s_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:k_ -s 400 -c 44 -f -o $O/r2s_step64 \
  python bench.py --steps 2 --warmup 3 --no-side-runs --no-cpu-baseline --no-e2e > $O/r2s_step64.log 2>&1; echo "full rc=$?"
ls -la $O/r2s_step64.ncu-rep
