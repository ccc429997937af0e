#!/bin/bash
# round-2 batch W (2 GPUs): multi-GPU tests, bench --gpus 2 under torchrun (fleet, strong scaling, sharded single sequence)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q > $O/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2w_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
   > $O/r2w_bench_n2.json 2> $O/r2w_bench_n2.err; echo "bench n2 rc=$?"
timeout 300 python profiles/run_ingest.py "" > $O/r2w_ingest.log 2>&1; cat $O/r2w_ingest.log | tail -2
