#!/bin/bash
# Build a variant of the working tree's library with extra nvcc flags:
#   bash scripts/ab_variant.sh <name> "<flags>"  ->  r-pcc_b200/build/ab/librpcc_<name>.so
set -e
NAME=$1; FLAGS=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
mkdir -p "$TMP/r-pcc_b200" "$ROOT/r-pcc_b200/build/ab"
cp -r "$ROOT/r-pcc_b200/csrc" "$TMP/r-pcc_b200/csrc"; cp -r "$ROOT/include" "$TMP/include"
make -C "$TMP/r-pcc_b200/csrc" -j 16 EXTRA="$FLAGS" > "$TMP/make.log" 2>&1 || { tail -20 "$TMP/make.log"; exit 1; }
cp "$TMP/r-pcc_b200/lib/librpcc_b200.so" "$ROOT/r-pcc_b200/build/ab/librpcc_$NAME.so"
rm -rf "$TMP"
echo "$ROOT/r-pcc_b200/build/ab/librpcc_$NAME.so"
