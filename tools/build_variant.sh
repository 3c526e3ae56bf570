#!/bin/bash
# Build an alternative libflowagg (same ABI) with extra nvcc flags for a same-box A/B:
#   tools/build_variant.sh x3 "-DFA_K1_EXP=3"   ->  netobserv_ebpf_agent_b200/libflowagg_x3.so  (load with FA_LIB_NAME=libflowagg_x3.so)
set -e
name=$1; flags=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
D=/tmp/fa_variant_$name; rm -rf $D; mkdir -p $D/csrc $D/include
cp $ROOT/netobserv_ebpf_agent_b200/csrc/*.cu $ROOT/netobserv_ebpf_agent_b200/csrc/*.cuh $ROOT/netobserv_ebpf_agent_b200/csrc/*.h $ROOT/netobserv_ebpf_agent_b200/csrc/Makefile $D/csrc/
cp $ROOT/include/flowagg.h $D/include/
cd $D/csrc
sed -i "s#\.\./\.\./include/flowagg.h#$D/include/flowagg.h#g" *.cu Makefile
sed -i "s#^OUT  = .*#OUT = $D/libflowagg_$name.so#" Makefile
make -j8 -s EXTRA="$flags"
cp $D/libflowagg_$name.so $ROOT/netobserv_ebpf_agent_b200/
grep -A3 "aggregate_kernelILb0ELb0ELb0E" aggregate.ptxas.log | grep -E "spill" || true
