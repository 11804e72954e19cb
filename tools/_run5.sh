set -x
mkdir -p gpurun_out; rm -f gpurun_out/mp_ab.txt
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_pytest.log 2>&1; tail -5 gpurun_out/r2e_pytest.log
timeout 300 python tools/mp_ab.py QAGNN_MP_WARPS=24 QAGNN_MP_WARPS=22 > gpurun_out/r2e_ab.log 2>&1; grep cfg2 gpurun_out/mp_ab.txt
timeout 400 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -3 gpurun_out/r2e_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cuda-graph > gpurun_out/r2e_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mp_headtile -s 2 -c 1 -o gpurun_out/r2e_mp -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2e_ncu_mp.log 2>&1
