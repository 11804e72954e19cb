set -x
mkdir -p gpurun_out
date
timeout 140 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2r_bench_n8.json 2> gpurun_out/r2r_bench_n8.err; echo "rc=$?"; date
tail -3 gpurun_out/r2r_bench_n8.err; cat gpurun_out/r2r_bench_n8.json | head -c 400
