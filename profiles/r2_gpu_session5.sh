set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29551 bench.py --gpus 8 --full > gpurun_out/r2_full_n8.json 2> gpurun_out/r2_full_n8.err; echo "full8 rc=$?"
tail -1 gpurun_out/r2_full_n8.json | cut -c1-900
timeout 600 $TR --master-port 29552 profiles/config5_check.py --golden > gpurun_out/r2_config5_golden_g8.log 2>&1; echo "golden8 rc=$?"
grep -E "CONFIG5|MISMATCH" gpurun_out/r2_config5_golden_g8.log
timeout 900 $TR --master-port 29553 profiles/config5_check.py --versus 360 --tables PK,PR,POmloop10 > gpurun_out/r2_config5_versus360_g8.log 2>&1; echo "versus8 rc=$?"
grep -E "CONFIG5|MISMATCH|sharded_vs" gpurun_out/r2_config5_versus360_g8.log | cut -c1-700
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/r2_nccl_%p.log timeout 900 $TR --master-port 29554 bench.py --gpus 8 --config5 --n5 600 > gpurun_out/r2_config5_n600_g8.json 2> gpurun_out/r2_config5_n600_g8.err; echo "config5 rc=$?"
tail -1 gpurun_out/r2_config5_n600_g8.json | cut -c1-2500
tail -3 gpurun_out/r2_config5_n600_g8.err
grep -h -i "NVLS\|via P2P" gpurun_out/r2_nccl_*.log | sort | uniq -c | sort -rn | head -6
ls gpurun_out/r2_nccl_*.log | head -1 | xargs -I{} cp {} gpurun_out/r2_nccl_rank_sample.log; rm -f gpurun_out/r2_nccl_[0-9]*.log
timeout 600 python - <<'PY' > gpurun_out/r2_multi_inprocess.log 2>&1
import sys, time, json
sys.path.insert(0, '.')
import bench, ccj_b200
par = 'params/rna_Turner04.par'
ctxs = [ccj_b200.Context(d, par, 2) for d in range(8)]
seqs = bench.workload2(1024)
ccj_b200.fold_batch_multi(ctxs, seqs[:64])
t0 = time.perf_counter(); folds = ccj_b200.fold_batch_multi(ctxs, seqs); dt8 = time.perf_counter() - t0
t0 = time.perf_counter(); one = ctxs[0].fold_batch(seqs); dt1 = time.perf_counter() - t0
gold = bench.load_goldens('folds_long.json', 'folds_config2.json')
chk = [(f, gold[s]) for s, f in zip(seqs, folds) if s in gold]
print(json.dumps({"check": "ccj_fold_batch_multi", "gpus": 8, "sequences": len(seqs), "seconds_8": dt8, "seconds_1": dt1,
                  "folds_per_s_8": len(seqs) / dt8, "folds_per_s_1": len(seqs) / dt1, "identical_to_one_gpu": folds == one,
                  "golden_checked": len(chk), "golden_parity": all(bench.same_as_golden(f, r) for f, r in chk)}))
PY
cat gpurun_out/r2_multi_inprocess.log | tail -2
