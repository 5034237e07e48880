for cfg in "1000 192" "64 64" "300 64" "8192 192"; do
NMX_DISABLE_CHAIN2T=1 timeout 300 python scripts/chain2t_check.py ref $cfg 2>&1 | tail -1
timeout 300 python scripts/chain2t_check.py cmp $cfg 2>&1 | tail -14
done
for m in 1 0; do echo "== NMX_DISABLE_CHAIN2T=$m"; if [ $m = 1 ]; then export NMX_DISABLE_CHAIN2T=1; else unset NMX_DISABLE_CHAIN2T; fi; timeout 300 python scripts/prof_step.py 8192 2>&1 | tail -13; done
