set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard.py "tests/test_gpu_parity.py::test_tuned_kernels_equal_generic_kernels" "tests/test_gpu_parity.py::test_beyond_the_tuned_range" "tests/test_gpu_parity.py::test_golden_benchmark_workloads" "tests/test_gpu_parity.py::test_beyond_the_reference_limit_against_the_cpu_restatement" -m gpu -x -q > gpurun_out/r2_gputest_s4.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2_gputest_s4.log
timeout 600 python - <<'PY' > gpurun_out/r2_shard_speed2.log 2>&1
import sys, time
sys.path.insert(0, '.')
import ccj_b200
from ccj_b200 import shard5
ctx = ccj_b200.Context(0, 'params/rna_Turner04.par', 2)
for n, world in [(150, 1), (200, 1), (200, 8), (260, 8), (300, 1)]:
    seq = shard5.config5_sequence(n)
    grp = shard5.LocalGroup(ctx, world)
    t0 = time.perf_counter()
    sh = grp.fold(seq)
    f = sh.traceback()
    print(n, world, grp.ms, 'wall', round(time.perf_counter() - t0, 2), 'tb_ms', sh.traceback_ms, f.energy, flush=True)
    grp.close()
PY
cat gpurun_out/r2_shard_speed2.log
