for d in 0 1 2 3 4 7; do echo "NMX_CHAIN2T_DBG=$d"; NMX_CHAIN2T_DBG=$d python scripts/fwd_train_time.py 8192 192 2>&1 | tail -1; done
NMX_DISABLE_CHAIN2T=1 python scripts/fwd_train_time.py 8192 192 2>&1 | tail -1
