#!/bin/bash
# Build variants of the product library for A/B timing inside ONE gpurun visit (boxes differ by 10 % in capped clock):
#   ab/lib_head.so   the last commit          ab/lib_<name>.so  the working tree with extra -D flags
# usage: tools/ab_build.sh name "-DFLAG=0" [name2 "-D..."]...
set -e
cd "$(dirname "$0")/.."
mkdir -p ab
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -shared"
SRC="oz_api.cu oz_scale.cu oz_gemm.cu oz_gemm_crt.cu oz_crt.cu oz_complex.cu oz_cxx_api.cu"
if [ "$1" = "--head" ]; then
  shift
  rm -rf /tmp/ab_head && git worktree prune && git worktree add -f /tmp/ab_head HEAD > /dev/null 2>&1
  (cd /tmp/ab_head && python tools/gen_tables.py > /dev/null && cd mixed-gemmul8_b200/csrc && S=""; for f in $SRC; do [ -f $f ] && S="$S $f"; done; nvcc $FLAGS -o /root/repo/ab/lib_head.so $S -ldl)
  git worktree remove --force /tmp/ab_head
fi
while [ $# -ge 2 ]; do
  (cd mixed-gemmul8_b200/csrc && nvcc $FLAGS $2 -o ../../ab/lib_$1.so $SRC -ldl)
  shift 2
done
ls -la ab
