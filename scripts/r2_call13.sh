timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_chain2_train_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r2_pair_fwd python scripts/fwd_train_time.py 8192 192 1 > gpurun_out/ncu_pair_fwd.log 2>&1; tail -2 gpurun_out/ncu_pair_fwd.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_chain2_bwd_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r2_pair_bwd python scripts/fwd_train_time.py 8192 192 1 > gpurun_out/ncu_pair_bwd.log 2>&1; tail -2 gpurun_out/ncu_pair_bwd.log
ls -la gpurun_out/*.ncu-rep
timeout 200 python -m pytest tests/test_mlp_gpu.py tests/test_training_gpu.py -m gpu -q 2>&1 | tail -3
