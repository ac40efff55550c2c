#!/bin/bash
# Variants of libccj_b200.so that differ in the -D flags of ccj_shard.cu (the lean sharded kernels):
#   profiles/build_shard_variants.sh name "-DFOO=1" ...  -> ccj_b200/variants/libccj_<name>.so (select with CCJ_B200_LIB)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
C=$ROOT/ccj_b200/csrc
V=$ROOT/ccj_b200/variants
mkdir -p $V/obj
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr -diag-suppress 20012 -I $ROOT/include"
PAR=("-DCCJ_PAR_TURNER04=\"$ROOT/params/rna_Turner04.par\"" "-DCCJ_PAR_DNA_MATHEWS04=\"$ROOT/params/dna_Matthews04.par\"")
for s in ccj_abi.cu ccj_kernels.cu ccj_peak.cu ccj_fill4.cu energy_model.cpp embedded_params.cpp; do
  o=$V/obj/${s%.*}.o
  nvcc $FLAGS "${PAR[@]}" -c -o $o $C/$s &
done
wait
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  ( nvcc $FLAGS $defs -c -o $V/obj/shard_$name.o $C/ccj_shard.cu &&
    nvcc -shared -o $V/libccj_$name.so $V/obj/ccj_abi.o $V/obj/ccj_kernels.o $V/obj/ccj_fill4.o $V/obj/ccj_peak.o $V/obj/shard_$name.o $V/obj/energy_model.o $V/obj/embedded_params.o -lcudart -ldl &&
    echo "built $name" ) &
  while [ $(jobs -r | wc -l) -ge 6 ]; do sleep 1; done
done
wait
