#!/bin/bash
# Builds variants of libccj_b200.so that differ only in the -D flags of ccj_fill4.cu (timing experiments):
#   profiles/build_variants.sh name1 "-DFOO=1" name2 "-DBAR=2 -DBAZ" ...
# -> ccj_b200/variants/libccj_<name>.so (git-ignored, travels to the GPU box); select with CCJ_B200_LIB.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
C=$ROOT/ccj_b200/csrc
V=$ROOT/ccj_b200/variants
mkdir -p $V/obj
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr -diag-suppress 20012 -I $ROOT/include"
PAR=("-DCCJ_PAR_TURNER04=\"$ROOT/params/rna_Turner04.par\"" "-DCCJ_PAR_DNA_MATHEWS04=\"$ROOT/params/dna_Matthews04.par\"")
for s in ccj_abi.cu ccj_kernels.cu ccj_peak.cu ccj_shard.cu energy_model.cpp embedded_params.cpp; do
  o=$V/obj/${s%.*}.o
  if [ ! -f $o ] || [ $C/$s -nt $o ] || [ -n "$(find $C -name '*.cuh' -newer $o -o -name '*.h' -newer $o)" ]; then
    nvcc $FLAGS "${PAR[@]}" -c -o $o $C/$s &
  fi
done
wait
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  ( { [ $V/obj/fill4_$name.o -nt $C/ccj_fill4.cu ] && [ "$(cat $V/obj/fill4_$name.defs 2>/dev/null)" == "$defs" ] || nvcc $FLAGS $defs -c -o $V/obj/fill4_$name.o $C/ccj_fill4.cu; } && echo "$defs" > $V/obj/fill4_$name.defs &&
    nvcc -shared -o $V/libccj_$name.so $V/obj/ccj_abi.o $V/obj/ccj_kernels.o $V/obj/fill4_$name.o $V/obj/ccj_peak.o $V/obj/ccj_shard.o $V/obj/energy_model.o $V/obj/embedded_params.o -lcudart -ldl &&
    echo "built $name" ) &
  # at most 4 at a time
  while [ $(jobs -r | wc -l) -ge 4 ]; do sleep 1; done
done
wait
