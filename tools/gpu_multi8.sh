#!/bin/bash
# (under gpurun --gpus N) short multi-GPU evidence run: peer-logging check + weak-scaling bench line at N and 1
N=${1:-8}; TAG=${2:-r2y}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    tools/dist_check.py 8192 20 command_direction 2>&1 | grep -E "PEER|MISMATCH|rank .* step|Error|error" | head -20 | tee gpurun_out/${TAG}_dist_check_${N}.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-cpu --no-e2e --no-sweep --no-configs > gpurun_out/${TAG}_bench_${N}.json 2> gpurun_out/${TAG}_bench_${N}.err
tail -c 300 gpurun_out/${TAG}_bench_${N}.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --no-sweep --no-configs > gpurun_out/${TAG}_bench_1.json 2>/dev/null
python - <<PY
import json
for n in ($N, 1):
    d=json.loads(open('gpurun_out/${TAG}_bench_%d.json' % n).read().strip().splitlines()[-1])
    print('N=%d value %.4g ms/step %.4f frac %.3f post %.1f us' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel']['kernel_us']))
    if d.get('strong_scaling'): print('  strong_scaling', json.dumps(d['strong_scaling'])[:400])
PY
