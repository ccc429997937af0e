#!/bin/bash
# round-2 batch Z: streaming kernels on high-priority context streams, ICP on a lowest-priority side stream
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
run() { # icp_prio stream_prio tag
  PTK_ICP_PRIORITY=$1 PTK_BENCH_STREAM_PRIORITY=$2 timeout 300 python bench.py --no-side-runs --no-cpu-baseline --no-e2e > $O/r2z_i$1s$2$3.json 2> $O/r2z_i$1s$2$3.err; echo "i$1 s$2 $3 rc=$?"
}
run 0 0 a
run 2 -1 a
run 0 -1 a
run 2 -1 b
run 0 0 b
