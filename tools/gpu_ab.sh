#!/bin/bash
# (under gpurun) A/B of two versions of csrc/post_kernel.cuh on ONE box (box-to-box variance is ~1 %):
#   bash tools/gpu_ab.sh <path of the alternative post_kernel.cuh inside the repo snapshot>
# runs the tree's version 3x, the alternative 3x, the tree's version 3x again (config 2, 1M envs).
ALT=${1:?path of the alternative post_kernel.cuh}
run() { for k in 1 2 3; do python bench.py --no-cpu --no-e2e --no-sweep --no-configs --steps 40 --warmup 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$1: step %.1f us post %.1f us action %.1f us' % (d['ms_per_step']*1e3, r['kernel']['kernel_us'], r['action_kernel']['kernel_us']))"; done; }
run tree
cp genesis_forge_b200/csrc/post_kernel.cuh /tmp/tree.cuh
cp "$ALT" genesis_forge_b200/csrc/post_kernel.cuh
bash tools/prep.sh > /dev/null 2>&1
run alternative
cp /tmp/tree.cuh genesis_forge_b200/csrc/post_kernel.cuh
bash tools/prep.sh > /dev/null 2>&1
run tree
