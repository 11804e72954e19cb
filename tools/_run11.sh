set -x
mkdir -p gpurun_out
date
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err; echo "rc=$?"; date
tail -3 gpurun_out/r2k_bench_n2.err; cat gpurun_out/r2k_bench_n2.json | head -c 300
