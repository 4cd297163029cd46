#!/bin/sh
# A/B build of the library with extra nvcc defines:  tests/build_variant.sh <tag> [-DNAME=VALUE ...]
# -> point-cloud-flow-matching_b200/lib/var/libpcfm_<tag>.so   (select with PCFM_B200_LIB=<path>)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
tag=$1; shift
out=$ROOT/point-cloud-flow-matching_b200/lib/var
mkdir -p $out/obj_$tag
for f in $ROOT/point-cloud-flow-matching_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  extra=""; [ $b = emd ] && extra="-ftz=true"
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
    --expt-relaxed-constexpr $extra "$@" -c $f -o $out/obj_$tag/$b.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o $out/libpcfm_$tag.so $out/obj_$tag/*.o -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC -lcudart
rm -rf $out/obj_$tag
echo $out/libpcfm_$tag.so
