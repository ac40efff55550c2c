mkdir -p gpurun_out
for w in 100000 112 96 80 64 48 32; do CCJ_WINLR_WIDE_FROM=$w python profiles/exp_win_width.py; done > gpurun_out/r2_exp_win_width.log 2>&1
cat gpurun_out/r2_exp_win_width.log
