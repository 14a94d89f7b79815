#!/bin/bash
# rebuild only the given translation units (default: force + jerk) and relink libbnmpc.so with the other objects as they are
# (for kernel experiments; `make` in csrc/ is the real build)
cd "$(dirname "$0")/../drone_attitude_control_b200/csrc"
TUS=${@:-model_force model_jerk}
for t in $TUS; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v $EXTRA -c $t.cu -o ../../build/csrc/$t.o 2> ../../build/csrc/$t.ptxas.log &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libbnmpc.so ../../build/csrc/*.o
grep -A3 "k_loop_step.*EdLi1" ../../build/csrc/model_force.ptxas.log ../../build/csrc/model_jerk.ptxas.log | grep -E "spill|registers" 
