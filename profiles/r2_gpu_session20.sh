mkdir -p gpurun_out
( for v in w4 w1 w8 w4b6 nowin; do echo "$v:"; CCJ_B200_LIB=$PWD/ccj_b200/variants/libccj_$v.so python profiles/shard_one.py 300 1 | tail -2; done ) > gpurun_out/r2_shard_lean_windows.log 2>&1; cat gpurun_out/r2_shard_lean_windows.log
