#!/bin/bash
# Builds scema_b200/libscema_hist_m4.so / _m6.so: the library with k_resample_pair compiled for 4 / 6 resident CTAs per SM
# (128 / 80 registers) instead of the default 5 (96), for an A/B on the GPU box (SCEMA_LIB=... python tools/k1_probe.py).
set -e
cd "$(dirname "$0")/../scema_b200/csrc"
make -s -j8
for m in 4 6; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-Wall \
        -DPR_MIN_CTAS=$m -c resample.cu -o build/resample_m$m.o
    objs=$(ls build/*.o | grep -v resample)
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libscema_hist_m$m.so $objs build/resample_m$m.o -lcudart_static -lpthread -ldl -lrt
done
ls -l ../libscema_hist*.so
