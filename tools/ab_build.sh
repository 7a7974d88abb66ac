#!/bin/bash
# Development aid: `tools/ab_build.sh NAME [extra nvcc flags]` compiles the current kernels into ab/NAME.so
# (git-ignored, travels with gpurun) so that tools/profile_run.py --lib ab/NAME.so can time several kernel
# variants on the same GPU box in one call.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p ab
make -s -C gama_tts_b200/csrc host_tables.o batch_plan.o model5_host.o events_host.o
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" \
  -c -o ab/$name.o gama_tts_b200/csrc/runtime.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ab/$name.so ab/$name.o \
  gama_tts_b200/csrc/host_tables.o gama_tts_b200/csrc/batch_plan.o gama_tts_b200/csrc/model5_host.o gama_tts_b200/csrc/events_host.o -cudart static
rm -f ab/$name.o
ls -la ab/$name.so
