python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29611 scripts/dp_check.py > gpurun_out/r2_dp_check.log 2>&1
echo "dp_check rc=$?"; grep -v "^\[W\|^$" gpurun_out/r2_dp_check.log | tail -40
