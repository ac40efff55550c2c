mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_s10.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_gputest_s10.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; echo "bench rc=$?"
python - <<'PY'
import json
b = json.load(open('gpurun_out/r2_bench_d.json'))
print(b['value'], b['e2e']['value'], b['ms_per_step'], b['roofline']['kernel_ms'], [ (r['fill_ms'], r['traceback_ms'], r['parity']) for r in b['config3']['runs']], b['config2']['folds_per_s'], b['config2']['parity'], b['golden_150nt'])
PY
