# 2-GPU DDPGfD training under torchrun: must reach the last eval AND exit on its own (captured NCCL graphs are released
# before destroy_process_group).  The outer timeout is the guard: rc 124 = hung at exit again.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  -m kinovagrasping_b200.train --envs 4096 --sweeps 10 --pretrain-updates 2000 --tau 0.01 > gpurun_out/r2o_train_2gpu.log 2>&1
rc=$?; t1=$(date +%s)
echo "train rc=$rc wall=$((t1-t0))s" | tee -a gpurun_out/r2o_train_2gpu.log
tail -12 gpurun_out/r2o_train_2gpu.log | cut -c1-220
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2o_bench_2gpu.json 2> gpurun_out/r2o_bench_2gpu.err; echo "bench2 rc=$?"; cut -c1-300 gpurun_out/r2o_bench_2gpu.json
