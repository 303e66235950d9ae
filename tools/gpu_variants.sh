#!/bin/bash
# Usage (under gpurun): bash tools/gpu_variants.sh <tag>   -- post-kernel timing under the GFB_DEBUG experiment switches
TAG=${1:-rX}
mkdir -p gpurun_out
for dbg in ${VARIANTS:-0 2 8}; do
  GFB_DEBUG=$dbg timeout 180 python bench.py --no-cpu --no-e2e --no-sweep --no-configs --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('GFB_DEBUG=$dbg step %.1f us post %.1f us action %.1f us observe %.1f us' % (d['ms_per_step']*1e3, r['kernel']['kernel_us'], r['action_kernel']['kernel_us'], r['small_kernels']['observe_kernel']['kernel_us']))"
done | tee gpurun_out/variants_$TAG.txt
