timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench.err
for c in C1 C4; do timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_$c.csv python bench.py --config $c --steps 3 --warmup 3 --no-graph > /dev/null 2>&1; python scripts/ncu_launches_summary.py gpurun_out/r2_launches_$c.csv > gpurun_out/r2_launches_${c}_summary.txt; tail -12 gpurun_out/r2_launches_${c}_summary.txt; done
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.3f'%d['ms_per_step'], 'e2e %.3f'%d['e2e']['ms_per_step'], 'rays/s %.0f'%d['value'], 'frac %.4f'%d['step_tensor_frac']['frac'], {k:round(v['ms_per_step'],3) for k,v in r['modes'].items()}, 'wgrad %.3f'%r['wgrad']['ms_per_step'], 'chain frac %.4f'%r['frac'], 'traffic', r['traffic'])
print('render', d['render']['value'], d['render']['tensor_frac_of_sustained'])
for k,v in d['configs'].items(): print(k, {a:b for a,b in v.items() if a not in ('workload',)})
print('cpu', d['cpu_baseline']); print('clocks', d['clocks']); print('launches', d['gpu_launches'])
P
