timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/r2_bench_v14.json 2> gpurun_out/r2_bench_v14.err; tail -2 gpurun_out/r2_bench_v14.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_v14.json').read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.3f'%d['ms_per_step'], 'e2e %.3f'%d['e2e']['ms_per_step'], 'rays/s %.0f'%d['value'], 'frac', d['step_tensor_frac']['frac'], {k:round(v['ms_per_step'],3) for k,v in r['modes'].items()}, 'wgrad %.3f'%r['wgrad']['ms_per_step'], 'chain frac', r['frac'], 'render', d['render']['value'])
P
timeout 300 python bench.py --rays-per-gpu 1024 --steps 20 --warmup 5 --no-render --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=1024 ms/step %.3f'%d['ms_per_step'])"
