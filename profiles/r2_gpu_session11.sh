mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29561 bench.py --gpus 8 --config5 --n5 600 > gpurun_out/r2_config5_n600_g8_b.json 2> gpurun_out/r2_config5_n600_g8_b.err; echo "config5 rc=$?"
tail -1 gpurun_out/r2_config5_n600_g8_b.json | cut -c1-1400
timeout 900 $TR --master-port 29562 profiles/config5_check.py --golden --versus 400 --tables PK,PR > gpurun_out/r2_config5_check_g8_b.log 2>&1; echo "check rc=$?"
grep -E "CONFIG5|MISMATCH|sharded_vs" gpurun_out/r2_config5_check_g8_b.log | cut -c1-600
timeout 600 $TR --master-port 29563 bench.py --gpus 8 --full > gpurun_out/r2_full_n8_b.json 2> gpurun_out/r2_full_n8_b.err; echo "full8 rc=$?"
tail -1 gpurun_out/r2_full_n8_b.json | cut -c1-400
timeout 600 $TR --master-port 29564 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench8 rc=$?"
tail -1 gpurun_out/r2_bench_n8.json | cut -c1-500
