#!/bin/bash
# Build an A/B variant of the library that differs only in the reactor rollout TU:
#   tools/ab_build.sh <name> <extra nvcc flags...>   ->  neorl-industrial-gym_b200/_ab/libnig_b200_<name>.so
# (run it with NIG_LIB_PATH=... ; the other objects are taken from the regular build directory)
set -e
cd "$(dirname "$0")/../neorl-industrial-gym_b200/csrc"
name=$1; shift
mkdir -p ../_ab
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden"
$NV "$@" -c -o ../_ab/nig_rollout_reactor_$name.o nig_rollout_reactor.cu
objs=$(ls ../build/*.o | grep -v nig_rollout_reactor.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -Xcompiler -fPIC -o ../_ab/libnig_b200_$name.so $objs ../_ab/nig_rollout_reactor_$name.o -ldl
echo built ../_ab/libnig_b200_$name.so
