mkdir -p gpurun_out
for cfg in "32 200000" "32 400000" "64 200000" "64 400000" "128 400000" "128 2000000"; do
  set -- $cfg
  GFB_DEBUG=8 GFB_NO_SPEC=1 GFB_TILE=$1 timeout 120 python tools/loop_stress.py $2 40 2>&1 | tail -1
done | tee gpurun_out/r2p_loop_stress.txt
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -vE "UserWarning|torch.tensor\(|^  warnings" | tail -15 | cut -c1-250) | tee gpurun_out/r2p_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-configs > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; tail -c 1500 gpurun_out/r2p_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench.json').read().strip().splitlines()[-1])
r=d['roofline']
print('step %.1f us frac %.3f post %.1f action %.1f small %s sweep %s' % (d['ms_per_step']*1e3, r['frac'], r['kernel']['kernel_us'], r['action_kernel']['kernel_us'], {k:(v['kernel_us'] if isinstance(v,dict) else v) for k,v in r['small_kernels'].items()}, {k:round(v['ms_per_step']*1e3,1) for k,v in d['sweep'].items()}))
PY
