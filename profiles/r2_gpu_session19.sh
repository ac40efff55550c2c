mkdir -p gpurun_out
timeout 300 python -m pytest "tests/test_gpu_shard.py::test_in_process_group_matches_reference_tables_and_folds" "tests/test_gpu_shard.py::test_in_process_group_on_a_benchmark_size_fold" "tests/test_gpu_parity.py::test_tuned_kernels_equal_generic_kernels" "tests/test_gpu_parity.py::test_beyond_the_tuned_range" -m gpu -x -q > gpurun_out/r2_gputest_s19.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_gputest_s19.log
( echo "default (8 blocks/SM, unroll 2):"; python profiles/shard_one.py 300 1
  for v in b6u2 b10u2 b12u2 b8u1 b8u4 b5u4 b12u1; do echo "$v:"; CCJ_B200_LIB=$PWD/ccj_b200/variants/libccj_$v.so python profiles/shard_one.py 300 1 | tail -1; done ) > gpurun_out/r2_shard_lean_variants.log 2>&1; cat gpurun_out/r2_shard_lean_variants.log
