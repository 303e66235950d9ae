#!/bin/bash
# (under gpurun) slab-size / loop experiments, one line per run
mkdir -p gpurun_out
one() {  # label, env..., -- bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --no-cpu --no-e2e --no-sweep --no-configs --steps 20 --warmup 5 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$label: step %.1f us post %.1f us (%.3f) action %.1f us small %s lib %s' % (d['ms_per_step']*1e3, r['kernel']['kernel_us'], r['kernel']['frac'], r['action_kernel']['kernel_us'], {k:round(v['kernel_us'],1) for k,v in r['small_kernels'].items() if isinstance(v,dict)}, d['kernel_variant']))" 2>&1 | tail -1
}
{
one "cd 1M tile128" X=1 --
one "cd 1M tile64" GFB_TILE=64 --
one "cd 1M tile32" GFB_TILE=32 --
one "cd 1M tile128 noloop" GFB_DEBUG=4 --
one "contacts 1M" X=1 -- --config contacts
one "contacts 1M noloop" GFB_DEBUG=4 -- --config contacts
one "contacts 1M tile128" GFB_TILE=128 -- --config contacts
one "humanoid 1M" X=1 -- --config berkeley_humanoid
one "humanoid 1M noloop" GFB_DEBUG=4 -- --config berkeley_humanoid
one "rough 262144" X=1 -- --config rough_terrain --num-envs 262144
one "rough 262144 noloop" GFB_DEBUG=4 -- --config rough_terrain --num-envs 262144
} | tee gpurun_out/r2q_exp1.txt
