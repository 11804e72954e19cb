set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err; tail -5 gpurun_out/r2j_bench_n2.err; cat gpurun_out/r2j_bench_n2.json | head -c 600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2j_ref_n2.json 2> gpurun_out/r2j_ref_n2.err; cat gpurun_out/r2j_ref_n2.json | head -c 400
