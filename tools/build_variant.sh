#!/usr/bin/env bash
# builds actinon_b200/variants/lib<name>.so with extra nvcc defines (kernel tuning experiments): tools/build_variant.sh name -DACN_...=..
set -euo pipefail
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p actinon_b200/variants build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-std=c++17 -O3 -use_fast_math -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function -Iinclude"
$NVCC $FLAGS "$@" -c actinon_b200/csrc/acn_tracer.cu -o build/acn_tracer_$name.o 2>/dev/null
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o actinon_b200/variants/lib$name.so build/acn_tracer_$name.o build/acn_model.o build/acn_host.o build/acn_interp.o build/acn_embed.o -cudart static -ldl
cuobjdump -res-usage actinon_b200/variants/lib$name.so 2>/dev/null | grep -A1 "IfLi2ELb0E" | grep -E "Function|REG" | paste - - | grep -E "k_direct|k_path|k_rays" | sed 's/Function _ZN3acn[0-9]\(k_[a-z]*\).*REG:\([0-9]*\) STACK:\([0-9]*\) SHARED:\([0-9]*\).*/  \1 REG \2 STACK \3 SHARED \4/' | tr '\n' ' '; echo " <- $name"
