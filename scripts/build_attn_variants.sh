#!/bin/bash
# Build A/B variants of libpi3b200.so that differ only in attn.cu compile-time knobs (same ABI), into
# pi3_slam_b200/build/variants/<name>.so; scripts/ab_attn.sh times them on the same box (PI3B_LIB selects the library).
# usage: scripts/build_attn_variants.sh "name:-DFLAG=.. -DFLAG2=.." ...      (name "r01" = the round-1 attn.cu from git)
set -e
cd "$(dirname "$0")/.."
python -m pi3_slam_b200.build > /dev/null
V=pi3_slam_b200/build/variants
mkdir -p $V
rm -f $V/*.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -O2 -I pi3_slam_b200/csrc"
OTHERS=$(ls pi3_slam_b200/build/*.o | grep -v "/attn.o")
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  (
  if [ "$name" = "r01" ]; then
    git show a1aa44c:pi3_slam_b200/csrc/attn.cu > $V/attn_r01.cu
    cat >> $V/attn_r01.cu <<'EOS'
extern "C" int pi3b_attn_fwd_hd64_lse(void*, const void*, const void*, const void*, int64_t, int64_t, int64_t, int, int, int, void*, int64_t, float*, int, int, int, int, float) { return -1; }
extern "C" int pi3b_attn_merge(void*, int, const void* const*, const float* const*, int64_t, void*, int64_t, float*, int, int) { return -1; }
EOS
    nvcc $FLAGS -c $V/attn_r01.cu -o $V/attn_$name.o 2>/dev/null
  else
    nvcc $FLAGS $defs -c pi3_slam_b200/csrc/attn.cu -o $V/attn_$name.o 2>/dev/null
  fi
  nvcc -shared --cudart shared -o $V/$name.so $V/attn_$name.o $OTHERS 2>/dev/null
  echo "built $V/$name.so  [$defs]"
  ) &
done
wait
rm -f $V/*.o $V/*.cu
