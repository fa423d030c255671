#!/bin/bash
# Build the -DSSAC_TRACE variant of the library (clock64 timelines) next to the product library.
set -e
cd "$(dirname "$0")/.."
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DSSAC_TRACE \
  -o super_sac_b200/libssac_b200_trace.so super_sac_b200/csrc/*.cu -L/usr/local/cuda/lib64/stubs -lcuda
