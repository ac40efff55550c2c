mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard.py "tests/test_gpu_parity.py::test_tuned_kernels_equal_generic_kernels" "tests/test_gpu_parity.py::test_beyond_the_tuned_range" "tests/test_gpu_parity.py::test_real_int16_wrap_matches_the_reference" "tests/test_gpu_parity.py::test_int16_negative_wrap_at_n213" tests/test_shell.py -m gpu -x -q > gpurun_out/r2_gputest_s9.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2_gputest_s9.log
for n in 150 200 300; do python profiles/shard_one.py $n 1; done > gpurun_out/r2_shard_speed4.log 2>&1; cat gpurun_out/r2_shard_speed4.log
CCJ_GENERIC_SCAN=1 python profiles/shard_one.py 200 1
