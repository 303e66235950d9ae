#!/bin/bash
mkdir -p gpurun_out
one() {  # label, env..., -- bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --no-cpu --no-e2e --no-sweep --no-configs --steps 30 --warmup 8 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$label: step %.1f us (%.3f) post %.1f us (%.3f) action %.1f us small %s %s' % (d['ms_per_step']*1e3, r['frac'], r['kernel']['kernel_us'], r['kernel']['frac'], r['action_kernel']['kernel_us'], {k:round(v['kernel_us'],1) for k,v in r['small_kernels'].items() if isinstance(v,dict)}, d['kernel_variant']['libraries']))" 2>&1 | tail -1
}
{
one "rough 262144 tile128" GFB_TILE=128 -- --config rough_terrain --num-envs 262144
one "rough 262144 tile64" GFB_TILE=64 -- --config rough_terrain --num-envs 262144
one "contacts 262144 tile128" GFB_TILE=128 -- --config contacts --num-envs 262144
one "contacts 262144 tile64" GFB_TILE=64 -- --config contacts --num-envs 262144
one "contacts 1M tile64" GFB_TILE=64 -- --config contacts
one "cd 262144 tile128" GFB_TILE=128 -- --num-envs 262144
one "cd 262144 tile64" GFB_TILE=64 -- --num-envs 262144
one "cd 65536 tile64" GFB_TILE=64 -- --num-envs 65536
one "cd 65536 tile128" GFB_TILE=128 -- --num-envs 65536
} | tee gpurun_out/r2ab_exp3.txt
