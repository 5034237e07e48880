python -m pytest tests -m gpu -x -q 2>&1 | tail -8
NMX_DISABLE_WGRAD_BATCH=1 python scripts/r2_dbg_wgrad.py ref > /dev/null; for p in 38605 1280 65536; do NMX_DISABLE_WGRAD_BATCH=1 python scripts/r2_dbg_wgrad.py ref $p >/dev/null 2>&1; python scripts/r2_dbg_wgrad.py cmp $p 2>&1 | awk '{print $NF, $1}' | sort -g | tail -2; done
python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29611 scripts/dp_check.py > gpurun_out/r2_dp_check.log 2>&1
echo "dp_check rc=$?"; grep -v "^\[W\|^$\|^W1\|^\*\*\*" gpurun_out/r2_dp_check.log | tail -12
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_v13.json 2> gpurun_out/r2_bench_v13.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_v13.err; cut -c1-300 gpurun_out/r2_bench_v13.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_v13_2gpu.json 2> gpurun_out/r2_bench_v13_2gpu.err; echo "bench2 rc=$?"; grep -v "^\[W\|^$\|^W1\|^\*\*\*" gpurun_out/r2_bench_v13_2gpu.err | tail -5; cut -c1-300 gpurun_out/r2_bench_v13_2gpu.json
