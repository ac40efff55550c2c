mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard.py "tests/test_gpu_parity.py::test_tuned_kernels_equal_generic_kernels" "tests/test_gpu_parity.py::test_beyond_the_tuned_range" "tests/test_gpu_parity.py::test_real_int16_wrap_matches_the_reference" -m gpu -x -q > gpurun_out/r2_gputest_s13.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_gputest_s13.log
for n in 200 300; do python profiles/shard_one.py $n 1; done > gpurun_out/r2_shard_speed5.log 2>&1; cat gpurun_out/r2_shard_speed5.log
timeout 600 python bench.py --full > gpurun_out/r2_full_n1_b.json 2> gpurun_out/r2_full_n1_b.err; echo "full rc=$?"
tail -1 gpurun_out/r2_full_n1_b.json | cut -c1-330
