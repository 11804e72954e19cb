set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2s_pytest.log 2>&1; tail -3 gpurun_out/r2s_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; tail -2 gpurun_out/r2s_smoke.log
timeout 400 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; tail -3 gpurun_out/r2s_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s_ref.json 2> gpurun_out/r2s_ref.err; cat gpurun_out/r2s_ref.json | head -c 500
