for cfg in "64 64" "1000 192"; do
NMX_DISABLE_CHAIN2T=1 NMX_DISABLE_CHAIN2B=1 timeout 120 python scripts/chain2t_check.py ref $cfg 2>&1 | tail -1
timeout 120 python scripts/chain2t_check.py cmp $cfg 2>&1 | tail -2
done
rm -f gpurun_out/*.pt
timeout 120 python scripts/fwd_train_time.py 8192 192 2>&1 | tail -1
timeout 120 python scripts/fwd_train_time.py 8192 64 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
