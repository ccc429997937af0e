#!/bin/bash
# round-2 batch AA: fleet workers outnumbering the host cores (8 contexts + main on 4 cores, as on an 8-GPU box with 32 cores):
# yielding poll against blocking-sync event
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
run() { # wait tag
  PTK_FLEET_WAIT=$1 timeout 300 taskset -c 0-3 python bench.py --no-side-runs --no-cpu-baseline > $O/r2aa_$1$2.json 2> $O/r2aa_$1$2.err; echo "$1 $2 rc=$?"
}
run yield a
run block a
run yield b
run block b
