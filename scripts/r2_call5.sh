python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29611 scripts/dp_check.py > gpurun_out/r2_dp_check.log 2>&1
echo "dp_check rc=$?"; grep -v "^\[W\|^$\|^W1\|^\*\*\*" gpurun_out/r2_dp_check.log | tail -6
for b in 1024 2048 4096 8192; do for m in 0 999999999999; do
NMX_WGRAD_BATCH_MAX_POINTS=$m python bench.py --rays-per-gpu $b --steps 20 --warmup 5 --no-render --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('B=$b batch_max=$m', 'ms/step %.3f'%d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in r['modes'].items()}, 'wgrad %.3f ms in %d launches'%(r['wgrad']['ms_per_step'], r['wgrad']['launches']))"
done; done
