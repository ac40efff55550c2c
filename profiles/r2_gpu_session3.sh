set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv
nvidia-smi topo -m | head -12
timeout 900 python -m pytest tests/test_gpu_shard.py -m gpu -x -q > gpurun_out/r2_gpushard_2gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2_gpushard_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 profiles/config5_check.py --versus 300 > gpurun_out/r2_config5_versus300_g2.log 2>&1; echo "versus rc=$?"
grep -v "^W1\|^\*\*\*" gpurun_out/r2_config5_versus300_g2.log | tail -8
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --config5 --n5 300 > gpurun_out/r2_config5_n300_g2.json 2> gpurun_out/r2_config5_n300_g2.err; echo "config5 rc=$?"
tail -1 gpurun_out/r2_config5_n300_g2.json | cut -c1-1500
grep -i "NVLS\|P2P\|via" gpurun_out/r2_config5_n300_g2.err | head -8
