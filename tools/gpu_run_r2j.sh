mkdir -p gpurun_out
timeout 600 python -m kinovagrasping_b200.train --envs 4096 --sweeps 20 --updates 100 --pretrain-updates 4000 --eval-every 5 --tau 0.01 > gpurun_out/r2j_train_1gpu.log 2>&1
grep -E "eval|pretrain" gpurun_out/r2j_train_1gpu.log | cut -c1-200; tail -2 gpurun_out/r2j_train_1gpu.log | cut -c1-300
