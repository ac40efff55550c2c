mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for n in ${SIZES:-600}; do
  timeout 400 $TR --master-port 2958$((n % 10)) bench.py --gpus $N --config5 --n5 $n > gpurun_out/r2_config5_lean_n${n}_g${N}.json 2> gpurun_out/r2_config5_lean_n${n}_g${N}.err; echo "n=$n rc=$?"
  tail -1 gpurun_out/r2_config5_lean_n${n}_g${N}.json | cut -c1-900
  tail -2 gpurun_out/r2_config5_lean_n${n}_g${N}.err | cut -c1-300
done
