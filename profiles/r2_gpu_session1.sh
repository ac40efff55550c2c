set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest.log
tail -5 gpurun_out/r2_gputest.log
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench.err
timeout 600 python bench.py --full > gpurun_out/r2_full_n1.json 2> gpurun_out/r2_full_n1.err; echo "full rc=$?"
cat gpurun_out/r2_full_n1.json | cut -c1-600
for tool in racecheck initcheck memcheck; do
  timeout 600 compute-sanitizer --tool $tool python profiles/sanitize_small.py > gpurun_out/r2_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
  tail -3 gpurun_out/r2_sanitizer_$tool.log
done
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_traffic.csv python profiles/traffic_workload.py 32 > gpurun_out/r2_traffic.log 2>&1; echo "ncu traffic rc=$?"
timeout 900 ncu --set full --metrics smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active --clock-control none --import-source on --launch-skip 390 --launch-count 14 -o gpurun_out/r2_levels_n150x32 python profiles/traffic_workload.py 32 > gpurun_out/r2_ncu_levels.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
