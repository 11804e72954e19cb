set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 3 -o gpurun_out/r2l_gemm -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-graph > gpurun_out/r2l_ncu_gemm.log 2>&1
