# 2-GPU bench line (torchrun, one rank per GPU) at HEAD
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2u_bench_2gpu.json 2> gpurun_out/r2u_bench_2gpu.err; echo "bench2 rc=$?"; cut -c1-300 gpurun_out/r2u_bench_2gpu.json
