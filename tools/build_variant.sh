#!/bin/bash
# usage: tools/build_variant.sh <tag> <extra nvcc flags...>   ->  ntm_tracker_b200/libntm_b200_<tag>.so
# A second build of the library with experiment macros, selected at run time with NTM_B200_LIB=<path> (development only).
set -e
tag=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
obj=$root/build/obj_$tag
mkdir -p "$obj"
pids=()
for cu in "$root"/ntm_tracker_b200/csrc/*.cu; do
  o=$obj/$(basename "${cu%.cu}").o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -split-compile 0 \
    "$@" -I"$root/include" -I"$root/ntm_tracker_b200/csrc" -c "$cu" -o "$o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
/usr/local/cuda/bin/nvcc -shared -o "$root/ntm_tracker_b200/libntm_b200_$tag.so" "$obj"/*.o 2>/dev/null
echo "built ntm_tracker_b200/libntm_b200_$tag.so"
