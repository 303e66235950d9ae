#!/bin/bash
mkdir -p gpurun_out
one() {  # label, env..., -- bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --no-cpu --no-e2e --no-sweep --no-configs --steps 30 --warmup 5 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$label: step %.1f us (%.3f) post %.1f us action %.1f us (%.3f) small %s' % (d['ms_per_step']*1e3, r['frac'], r['kernel']['kernel_us'], r['action_kernel']['kernel_us'], r['action_kernel']['frac'], {k:round(v['kernel_us'],1) for k,v in r['small_kernels'].items() if isinstance(v,dict)}))" 2>&1 | tail -1
}
{
one "action tile 128" X=1 --
one "action tile 64" GFB_ACTION_TILE=64 --
one "action tile 256" GFB_ACTION_TILE=256 --
one "action tile 32" GFB_ACTION_TILE=32 --
} | tee gpurun_out/r2aa_exp2.txt
