# 1/2/4/8-GPU scaling table on ONE box (weak headline + strong block + sharded render + dp_check per N)
python -m pytest tests/test_dp_gpu.py -m gpu -x -q 2>&1 | tail -3
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/r2_scale_n$n.json 2> gpurun_out/r2_scale_n$n.err; echo "N=$n rc=$?"
done
python bench.py --gpus 1 --steps 20 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err; echo "N=1 rc=$?"
python - <<'P'
import json
for n in (1,2,4,8):
    try:
        d=json.loads(open(f'gpurun_out/r2_scale_n{n}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(n, 'no line', e); continue
    s=d.get('strong') or {}; r=d.get('render') or {}
    print(f"N={n} weak {d['value']/1e6:.3f} M rays/s {d['ms_per_step']:.3f} ms | e2e {d['e2e']['ms_per_step']:.3f} ms | strong {s.get('ms_per_step')} ms x{s.get('speedup_vs_n1_same_run')} | render {r.get('value')} Msamples/s {r.get('ms_per_frame')} ms | dp {d.get('dp_check')} | clocks {d['clocks']}")
P
