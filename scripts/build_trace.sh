#!/bin/bash
# development aid: the kernel library with device-side timeline stamps compiled in (-DHULC_RNN_TRACE -DHULC_TC_TRACE)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/trace hulc_b200/lib
for f in hulc_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr \
    -DHULC_RNN_TRACE -DHULC_TC_TRACE -I hulc_b200/csrc -I include -c "$f" -o build/trace/$(basename "$f" .cu).o 2>/dev/null &
done
wait
nvcc -shared -o hulc_b200/lib/libhulc_trace.so build/trace/*.o -lcudart -lcuda
echo hulc_b200/lib/libhulc_trace.so
