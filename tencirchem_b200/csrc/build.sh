#!/usr/bin/env bash
# Build libtcc_b200.so for sm_100a in-tree (the .so travels to the GPU box with the repo snapshot).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../lib"
mkdir -p "$out"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
    -Xcompiler -fPIC,-O3,-Wall,-pthread -shared ${TCC_NVCC_EXTRA:-} \
    -o "$out/${TCC_LIB_NAME:-libtcc_b200.so}" "$here/engine.cu" "$here/kernels.cu" "$here/batch_kernels.cu" "$here/rdm_kernels.cu" "$here/plan.cpp"
echo "built $out/${TCC_LIB_NAME:-libtcc_b200.so}"
