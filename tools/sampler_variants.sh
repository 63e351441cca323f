#!/bin/bash
# Builds sampler.cu with different tuning knobs into tools/variants/*.so (run HERE), prints the list.
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/variants
rm -f tools/variants/*.so
CS=dynamic_multiview_3d_b200/csrc
while read -r name flags; do
  [ -z "$name" ] && continue
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr $flags -shared -o tools/variants/$name.so $CS/sampler.cu $CS/api.cu &
done <<LIST
a_tile -DSAMPLER_TMA=0
tma_40_48_f5 -DSAMPLER_BOX=40 -DSAMPLER_BOX_L=48 -DSAMPLER_TMA_MINBLOCKS_FWD3=5
tma_36_44_f5 -DSAMPLER_TMA_MINBLOCKS_FWD3=5
tma_36_44_f6
tma_40_44_f6 -DSAMPLER_BOX=40
tma_36_44_f7 -DSAMPLER_TMA_MINBLOCKS_FWD3=7
tma_28_44_f6 -DSAMPLER_BOX=28
LIST
wait
ls tools/variants
