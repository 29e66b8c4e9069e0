#!/bin/bash
# A/B builds of the attention kernel only (profiling): gpurun_variants/libdav2_attn_<name>.so = the shipped objects with
# attention.cu recompiled under extra macros.  Usage: scripts/build_attn_variants.sh name1:"-DFLAG ..." name2:...
# (name "r1" = the round-1 kernel from git)
set -e
cd "$(dirname "$0")/../enhanced-3d-reconstruction-in-colonoscopy-using-monocular-depth-and-pose-estimation_b200/csrc"
make -j16 >/dev/null
mkdir -p ../../gpurun_variants build_var
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
OTHER=$(ls build/*.o | grep -v attention.o)
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  src=attention.cu
  if [ "$name" = "r1" ]; then git show f161b6f:./attention.cu | sed 's/first_use_on_device(&tag)/true/' > build_var/attention_r1.cu; src=build_var/attention_r1.cu; defs="-I."; fi
  nvcc $FLAGS $defs -c $src -o build_var/attention_$name.o &
done
wait
for spec in "$@"; do
  name=${spec%%:*}
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../gpurun_variants/libdav2_attn_$name.so $OTHER build_var/attention_$name.o -lcudart
done
ls -la ../../gpurun_variants/
