set -x
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_v15_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-render --no-configs --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python scripts/ncu_launches_summary.py gpurun_out/r2_v15_launches.csv > gpurun_out/r2_v15_launches_summary.txt; head -30 gpurun_out/r2_v15_launches_summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mlp_chain --launch-skip 15 --launch-count 5 -o gpurun_out/r2_v15_chain python scripts/prof_step.py 8192 > gpurun_out/ncu_chain.log 2>&1; tail -2 gpurun_out/ncu_chain.log
timeout 400 ncu --set full --clock-control none -k regex:wgrad --launch-skip 54 --launch-count 18 -o gpurun_out/r2_v15_wgrad python scripts/prof_step.py 8192 > gpurun_out/ncu_wgrad.log 2>&1; tail -2 gpurun_out/ncu_wgrad.log
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct --clock-control none -k regex:"composite|sample_pdf|sort_merge|hashgrid|pe_|gen_rays|sh_encode|assemble" -c 400 --csv --log-file gpurun_out/r2_aux_ncu.csv python scripts/bench_aux.py > gpurun_out/r2_aux_events.json 2> gpurun_out/ncu_aux.err; tail -3 gpurun_out/ncu_aux.err
ls -la gpurun_out/*.ncu-rep
