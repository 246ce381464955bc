#!/bin/bash
# Build a tuning variant of the library next to the production one:
#   profiles/build_variant.sh tag "-DPGW_TMA_L2_AHEAD=3 ..."   ->  scratch/lib_tag.so
# and run it with PGW_B200_LIB=scratch/lib_tag.so python bench.py ...
set -e
tag=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
d=$root/scratch/build_$tag
mkdir -p $d
for f in pgw_timestep pgw_column_tma pgw_staged pgw_ops pgw_step02 pgw_nanterp pgw_misc; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I$root/include "$@" \
       -c $root/pgw4era5_b200/csrc/$f.cu -o $d/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $root/scratch/lib_$tag.so $d/*.o
echo $root/scratch/lib_$tag.so
