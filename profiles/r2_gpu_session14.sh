mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29571 bench.py --gpus 8 --config5 --n5 600 > gpurun_out/r2_config5_n600_g8_c.json 2> gpurun_out/r2_config5_n600_g8_c.err; echo "config5 rc=$?"
tail -1 gpurun_out/r2_config5_n600_g8_c.json | cut -c1-900
timeout 600 $TR --master-port 29572 profiles/config5_check.py --golden --versus 400 --tables PK,PR > gpurun_out/r2_config5_check_g8_c.log 2>&1; echo "check rc=$?"
grep -E "CONFIG5|MISMATCH|sharded_vs" gpurun_out/r2_config5_check_g8_c.log | cut -c1-500
timeout 300 python -m pytest tests/test_gpu_shard.py::test_in_process_nccl_group_over_all_gpus -m gpu -x -q 2>&1 | tail -2
