mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard.py tests/test_cli.py "tests/test_gpu_parity.py::test_other_models_against_reference_binary_on_the_box" "tests/test_gpu_parity.py::test_in_process_multi_context_dealing" -m gpu -x -q > gpurun_out/r2_gputest_s12.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2_gputest_s12.log
