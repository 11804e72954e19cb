set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "stress or large_graph or kept_switch or pool or decoder" > gpurun_out/r2i_pytest.log 2>&1; tail -3 gpurun_out/r2i_pytest.log
timeout 600 python tools/bench_configs.py cfg5 > gpurun_out/r2i_cfg5.jsonl 2> gpurun_out/r2i_cfg5.err; cat gpurun_out/r2i_cfg5.jsonl; tail -3 gpurun_out/r2i_cfg5.err
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -3 gpurun_out/r2i_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2i_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cuda-graph > gpurun_out/r2i_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mp_slice -c 3 -o gpurun_out/r2i_cfg5 -f python tools/bench_configs.py cfg5 > gpurun_out/r2i_ncu_cfg5.log 2>&1
