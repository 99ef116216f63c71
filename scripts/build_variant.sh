#!/bin/bash
# usage: scripts/build_variant.sh <name> [nvcc -D flags...]  ->  variants/libhsr_<name>.so (A/B timing with scripts/variant_bench.py)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
mkdir -p "$ROOT/variants"
rm -rf /tmp/obj_$name
make -C "$ROOT/hypersonic-rans_b200" -j8 OBJ=/tmp/obj_$name LIB="$ROOT/variants/libhsr_$name.so" EXTRA="$*" > /tmp/make_$name.log 2>&1 || { tail -30 /tmp/make_$name.log; exit 1; }
echo "$name built"
