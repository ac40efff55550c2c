mkdir -p gpurun_out
python profiles/exp_wave_size.py > gpurun_out/r2_exp_wave_size.log 2>&1; tail -20 gpurun_out/r2_exp_wave_size.log | head -17
for n in 150 200 300; do python profiles/shard_one.py $n 1; done > gpurun_out/r2_shard_speed3.log 2>&1; cat gpurun_out/r2_shard_speed3.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_4d_shard --launch-skip 100 --launch-count 1 -o gpurun_out/r2_k4dshard_n200 python profiles/shard_one.py 200 1 > gpurun_out/r2_ncu_k4dshard.log 2>&1; tail -3 gpurun_out/r2_ncu_k4dshard.log
