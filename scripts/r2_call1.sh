set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/bw_probe.py > gpurun_out/r2_bw_probe_events.txt 2>&1; tail -20 gpurun_out/r2_bw_probe_events.txt
ncu --metrics dram__bytes_write.sum.per_second,dram__bytes_read.sum.per_second,dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_bw_probe_ncu.csv python scripts/bw_probe.py --once > gpurun_out/r2_bw_probe_ncu.log 2>&1
for b in 1024 2048 8192; do python scripts/prof_step.py $b > gpurun_out/r2_prof_step_$b.txt 2>&1; tail -14 gpurun_out/r2_prof_step_$b.txt; done
for b in 1024 2048 4096; do python bench.py --rays-per-gpu $b --steps 20 --warmup 5 --no-render --no-cpu-baseline > gpurun_out/r2_bench_b$b.json 2>gpurun_out/r2_bench_b$b.err; cat gpurun_out/r2_bench_b$b.json | cut -c1-400; done
