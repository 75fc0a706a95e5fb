#!/bin/bash
# Build an A/B variant of the library that differs in ONE translation unit:
#   tools/ab_build.sh <name> <tu: nig_rollout_reactor|nig_step|...> <extra nvcc flags...>
#   -> neorl-industrial-gym_b200/_ab/libnig_b200_<name>.so   (run with NIG_LIB_PATH=...; other objects come from build/)
set -e
cd "$(dirname "$0")/../neorl-industrial-gym_b200/csrc"
name=$1; tu=$2; shift 2
mkdir -p ../_ab
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden"
$NV "$@" -c -o ../_ab/${tu}_$name.o $tu.cu
objs=$(ls ../build/*.o | grep -v /$tu.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -Xcompiler -fPIC -o ../_ab/libnig_b200_$name.so $objs ../_ab/${tu}_$name.o -ldl
echo built ../_ab/libnig_b200_$name.so
