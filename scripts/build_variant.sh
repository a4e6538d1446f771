#!/bin/bash
# Build a variant of liboptp_b200.so with extra nvcc flags for ONE translation unit (A/B runs of kernel variants):
#   scripts/build_variant.sh <name> <unit.cu | path/to/alt_source.cu:unit.cu> [nvcc flags...]
# -> optpricer_b200/_lib/variants/liboptp_<name>.so   (load it with OPTP_LIB=<path>, see scripts/crr_ab.py)
set -e
PKG="$(cd "$(dirname "$0")/.." && pwd)/options-pricing-engine-black-scholes-binomial-monte-carlo_b200"
name="$1"; spec="$2"; shift 2
src="${spec%%:*}"; unit="${spec##*:}"
[ "$src" = "$unit" ] && src="$PKG/csrc/$unit"
out="$PKG/optpricer_b200/_lib/variants"; mkdir -p "$out" "$PKG/build/variants"
obj="$PKG/build/variants/${name}_${unit%.cu}.o"
extra=""; case "$unit" in mc_fp64.cu|dupire_table.cu) extra="-fmad=false";; esac
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --extended-lambda -Xcompiler -fPIC $extra "$@" \
     -I"$PKG/csrc" -c "$src" -o "$obj"
objs=""
for u in optp_api mc_host mc_fp32 mc_fp64 crr_batch dupire_table microbench rng_check comm; do
  if [ "$u.cu" = "$unit" ]; then objs="$objs $obj"; else objs="$objs $PKG/build/$u.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out/liboptp_$name.so" $objs -Xcompiler -fPIC -ldl
echo "$out/liboptp_$name.so"
