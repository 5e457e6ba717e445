#!/usr/bin/env bash
# builds a variant of the library: tools/exp_build.sh <name> <extra nvcc flags...>  -> actinon_b200/variants/lib<name>.so
set -euo pipefail
cd "$(dirname "$0")/.."
NAME=$1; shift
mkdir -p actinon_b200/variants build
[ -f build/acn_model.o ] || bash build.sh >/dev/null
nvcc -std=c++17 -O3 -use_fast_math -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Iinclude "$@" -c actinon_b200/csrc/acn_tracer.cu -o build/acn_tracer_$NAME.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o actinon_b200/variants/lib$NAME.so build/acn_tracer_$NAME.o build/acn_model.o build/acn_host.o build/acn_interp.o build/acn_embed.o -cudart static -ldl
echo "built variant $NAME: $*"
