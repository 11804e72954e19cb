mkdir -p gpurun_out
timeout 120 tools/bin/microbench > gpurun_out/r2w_microbench.txt 2>&1; cat gpurun_out/r2w_microbench.txt
