#!/bin/bash
# Development: build libraries that differ from lib/libfcb200.so in the shape macros of ONE source file, for A/B runs on the GPU:
#   tools/shape_variants.sh ops_elementwise.cu v1 "-DFCB_WC_U=1 -DFCB_WC_MB=4"   ->  tools/probes/libfcb200_v1.so
#   FCB200_LIB=tools/probes/libfcb200_v1.so python tools/opbench.py --ops windCooling
SRC=$1; TAG=$2; DEFS=$3
HERE=$(cd "$(dirname "$0")/.." && pwd)
P=$HERE/mi-fieldcalc_b200
OBJS=$(ls $P/build/*.o | grep -v "/$(basename $SRC .cu).o")
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -fvisibility=default $DEFS \
  -I $HERE/include -c $P/csrc/$SRC -o /tmp/variant_$TAG.o && \
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $HERE/tools/probes/libfcb200_$TAG.so /tmp/variant_$TAG.o $OBJS -lcudart_static -ldl -lrt -lpthread && echo built $TAG
