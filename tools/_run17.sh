set -x
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -3 gpurun_out/r2q_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cuda-graph > gpurun_out/r2q_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mp_headtile -s 2 -c 1 -o gpurun_out/r2q_mp -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2q_ncu_mp.log 2>&1
timeout 300 python tools/bench_configs.py cfg1 cfg3 > gpurun_out/r2q_configs.jsonl 2> gpurun_out/r2q_configs.err; cat gpurun_out/r2q_configs.jsonl; tail -3 gpurun_out/r2q_configs.err
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -c 60 --csv --log-file gpurun_out/r2q_cfg1_launches.csv python tools/bench_configs.py cfg1 > gpurun_out/r2q_ncu_cfg1.log 2>&1
