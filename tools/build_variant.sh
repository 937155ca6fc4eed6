#!/bin/bash
# usage: variant.sh name "-DFOO=1 ..."   -> voicebridge_b200/libvbgpu_<name>.so
set -e
cd /root/repo/voicebridge_b200/csrc
name=$1; shift
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function $@ -c score_tc.cu -o build/score_tc_$name.o
objs=$(ls build/*.o | grep -v score_tc)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libvbgpu_$name.so $objs build/score_tc_$name.o -lcudart_static -ldl -lpthread -lrt
rm build/score_tc_$name.o
