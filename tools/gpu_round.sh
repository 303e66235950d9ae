#!/bin/bash
# (under gpurun) GPU tests + the experiment list of the main configs; args: <tag>
TAG=${1:-r2x}
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || (timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -vE "UserWarning|torch.tensor\(|^  warnings" | tail -12 | cut -c1-250) | tee gpurun_out/${TAG}_tests.log
one() {  # label, env..., -- bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --no-cpu --no-e2e --no-configs --steps 20 --warmup 5 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$label: step %.1f us (%.3f) post %.1f us (%.3f) action %.1f us small %s sweep %s' % (d['ms_per_step']*1e3, r['frac'], r['kernel']['kernel_us'], r['kernel']['frac'], r['action_kernel']['kernel_us'], {k:round(v['kernel_us'],1) for k,v in r['small_kernels'].items() if isinstance(v,dict)}, {k:round(v['ms_per_step']*1e3,1) for k,v in d.get('sweep',{}).items()}))" 2>&1 | tail -1
}
{
one "cd 1M" X=1 --
one "contacts 1M" X=1 -- --config contacts --no-sweep
one "humanoid 1M" X=1 -- --config berkeley_humanoid --no-sweep
one "rough 262144" X=1 -- --config rough_terrain --num-envs 262144 --no-sweep
} | tee gpurun_out/${TAG}_exp.txt
