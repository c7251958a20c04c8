#!/bin/bash
# Build the C-ABI shared library for sm_100a (B200), in-tree.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
    -Xcompiler -fPIC -Xptxas -v -shared -I../../include \
    -o libvqe_b200.so vqe_b200.cu 2>&1 | grep -E "error|warning|registers|Compiling entry|spill" || true
ls -la libvqe_b200.so
