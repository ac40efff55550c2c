mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_s7.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2_gputest_s7.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench_c.err
python __graft_entry__.py --smoke 2>&1 | tail -2
nproc
timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "ref rc=$?"
cat gpurun_out/r2_bench_reference_arm.json | cut -c1-1500
