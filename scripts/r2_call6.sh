python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25
cat gpurun_out/r2_fullsize_parity.json
free -g | head -2; nproc
