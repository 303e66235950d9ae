#!/bin/bash
# (under gpurun --gpus N) multi-GPU evidence: peer-memory logging check, the 2-GPU tests, bench at N ranks
N=${1:-2}; TAG=${2:-r2w}
mkdir -p gpurun_out
nvidia-smi -L | head -8
for cfg in command_direction berkeley_humanoid; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    tools/dist_check.py 8192 30 $cfg 2>&1 | grep -E "PEER|MISMATCH|rank .* step|Error|error" | head -20
done | tee gpurun_out/${TAG}_dist_check_${N}.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3 | tee -a gpurun_out/${TAG}_dist_check_${N}.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-e2e --no-sweep > gpurun_out/${TAG}_bench_${N}.json 2> gpurun_out/${TAG}_bench_${N}.err
tail -c 600 gpurun_out/${TAG}_bench_${N}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_${N}.json').read().strip().splitlines()[-1])
print('N=%d value %.3g ms/step %.4f frac %.3f' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac']))
print('strong_scaling', json.dumps(d.get('strong_scaling'))[:600])
print('configs', {k:(round(v['ms_per_step']*1e3,1)) for k,v in d.get('configs',{}).items()})
PY
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-sweep --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1 same box: ms/step %.4f' % d['ms_per_step'])"
