mkdir -p gpurun_out
python profiles/exp_ablate.py --only-profiled base a1 a2 a4 a8 a16 a32 a65 a128 a255 2>&1 | tee gpurun_out/r2_ablate_final.log
