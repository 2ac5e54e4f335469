# 2-GPU bench line (torchrun, one rank per GPU) at HEAD, own arm and the CPU arm launched the same way
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2ah_bench_2gpu.json 2> gpurun_out/r2ah_bench_2gpu.err; echo "bench2 rc=$?"; cut -c1-300 gpurun_out/r2ah_bench_2gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2ah_bench_2gpu_ref.json 2>> gpurun_out/r2ah_bench_2gpu.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2ah_bench_2gpu_ref.json
