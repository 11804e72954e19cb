set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_pytest.log 2>&1; tail -5 gpurun_out/r2g_pytest.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -3 gpurun_out/r2g_bench.err
timeout 600 python tools/bench_configs.py cfg5 cfg5basic > gpurun_out/r2g_cfg5.jsonl 2> gpurun_out/r2g_cfg5.err; cat gpurun_out/r2g_cfg5.jsonl; tail -3 gpurun_out/r2g_cfg5.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mp_slice -c 3 -o gpurun_out/r2g_cfg5 -f python tools/bench_configs.py cfg5 > gpurun_out/r2g_ncu_cfg5.log 2>&1
