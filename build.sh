#!/usr/bin/env bash
# Builds libactinon_b200.so (CUDA tracer + host scene API) in-tree for sm_100a.
set -euo pipefail
cd "$(dirname "$0")"
SRC=actinon_b200/csrc
OUT=actinon_b200/libactinon_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-std=c++17 -O3 -use_fast_math -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function -Iinclude"
mkdir -p build
python tools/embed_src.py build/acn_embed.cpp      # device sources for the run-time (NVRTC) specialisation
pids=()
$NVCC $FLAGS -c $SRC/acn_tracer.cu -o build/acn_tracer.o & pids+=($!)
g++ -std=c++17 -O2 -fPIC -Wall -Iinclude -c $SRC/acn_model.cpp -o build/acn_model.o & pids+=($!)
g++ -std=c++17 -O2 -fPIC -Wall -Iinclude -c $SRC/acn_host.cpp -o build/acn_host.o & pids+=($!)
g++ -std=c++17 -O2 -fPIC -Wall -Iinclude -c $SRC/acn_interp.cpp -o build/acn_interp.o & pids+=($!)
g++ -std=c++17 -O1 -fPIC -c build/acn_embed.cpp -o build/acn_embed.o & pids+=($!)
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT build/acn_tracer.o build/acn_model.o build/acn_host.o build/acn_interp.o build/acn_embed.o -cudart static -ldl
echo "built $OUT"
