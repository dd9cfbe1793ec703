#!/bin/bash
# Builds of the library that differ in compile-time switches of the RNEA / ABA kernels, for scripts/gpu_ab.py:
#   scripts/build_variants.sh name "-DMB_PLAIN_RNEA=4 -DMB_PLAIN_ABA=0" [name2 "flags2" ...]
# Only the translation units that see the switches are recompiled (the plain RNEA / ABA instantiations and the flattener); the
# rest is taken from mecano_b200/csrc/build.  Output: mecano_b200/variants/<name>.so (git-ignored; travels to the GPU box).
set -e
cd "$(dirname "$0")/../mecano_b200/csrc"
mkdir -p ../variants build/var
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -ge 2 ]; do
   NAME=$1; FLAGS=$2; shift 2
   for g in 0 2; do
      nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $FLAGS -DMB_INST=$g -c kern_inst.cu -o build/var/${NAME}_kinst_$g.o &
   done
   CXXDEFS=$(echo " $FLAGS" | grep -o ' -D[^ ]*' | tr '\n' ' ') # only the -D switches go to the host compiler
   g++ -O2 -std=c++17 -fPIC -fvisibility=hidden $CXXDEFS -c flatten.cpp -o build/var/${NAME}_flatten.o &
   wait
   OTHERS=$(ls build/*.o | grep -v "build/kinst_0.o\|build/kinst_2.o\|build/flatten.o\|build/kernels_")
   nvcc $ARCH -shared -o ../variants/$NAME.so $OTHERS build/var/${NAME}_kinst_0.o build/var/${NAME}_kinst_2.o build/var/${NAME}_flatten.o -ldl
   echo "built ../variants/$NAME.so"
done
