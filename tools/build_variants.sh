#!/bin/bash
# Builds experiment variants of librt_b200 (different -D flags for the trace TUs) into lib/ as librt_b200_<tag>.so
set -e
cd "$(dirname "$0")/../metal4_raytracing_b200/csrc"
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -fmad=false -Xcompiler -fPIC -ccbin /usr/bin/g++ --expt-relaxed-constexpr -Xptxas -v"
make -j6 >/dev/null 2>&1
build() { # tag, defines...
  tag=$1; shift
  mkdir -p /tmp/variants/$tag
  for f in trace trace_wavefront selftest; do # one object for both halves of trace_wavefront.cu
    $NVCC $FLAGS "$@" -c $f.cu -o /tmp/variants/$tag/$f.o 2> /tmp/variants/$tag/$f.log &
  done
  wait
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/librt_b200_$tag.so rt_api.o bvh_build.o skin.o tiles.o post.o intersect.o host/renderer.o /tmp/variants/$tag/trace.o /tmp/variants/$tag/trace_wavefront.o /tmp/variants/$tag/selftest.o -cudart static
  echo "$tag: $(grep -A2 'k_wf_traverseILi8' /tmp/variants/$tag/trace_wavefront.log | grep -o 'Used [0-9]* registers' | head -1) $(grep -A1 'k_wf_traverseILi8' /tmp/variants/$tag/trace_wavefront.log | grep -o '[0-9]* bytes spill stores' | head -1)"
}
for spec in "$@"; do
  tag=${spec%%:*}; defs=${spec#*:}
  build $tag $defs
done
