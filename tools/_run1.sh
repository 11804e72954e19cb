set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 300 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2a_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mp_headtile -s 2 -c 1 -o gpurun_out/r2a_mp -f python bench.py --steps 2 --warmup 1 > gpurun_out/r2a_ncu_mp.log 2>&1
tail -3 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_bench.json
