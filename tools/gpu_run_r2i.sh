mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2i_bench_2gpu.json 2> gpurun_out/r2i_bench_2gpu.err
tail -c 1200 gpurun_out/r2i_bench_2gpu.json; tail -5 gpurun_out/r2i_bench_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 -m kinovagrasping_b200.train --envs 8192 --sweeps 6 --updates 100 --pretrain-updates 400 --eval-every 3 --tau 0.01 > gpurun_out/r2i_train_2gpu.log 2>&1
tail -12 gpurun_out/r2i_train_2gpu.log
