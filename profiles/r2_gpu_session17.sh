mkdir -p gpurun_out
python profiles/exp_ablate.py base prmt p8 k4u1b7 k4u1 k4u2 2>&1 | tee gpurun_out/r2_variants_roles.log
