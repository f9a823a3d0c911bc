#!/bin/bash
# A/B builds: compile the library of another revision next to the working tree's, for bit-for-bit
# comparisons of a kernel rewrite on the GPU box (RPCC_B200_LIB selects the library, _lib.py).
#   bash scripts/ab_build.sh <git-rev>   ->  r-pcc_b200/build/ab/librpcc_<rev>.so
set -e
REV=${1:-HEAD}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
git -C "$ROOT" archive "$REV" r-pcc_b200/csrc include | tar -x -C "$TMP"
make -C "$TMP/r-pcc_b200/csrc" -j 16 > "$TMP/make.log" 2>&1 || { tail -20 "$TMP/make.log"; exit 1; }
mkdir -p "$ROOT/r-pcc_b200/build/ab"
cp "$TMP/r-pcc_b200/lib/librpcc_b200.so" "$ROOT/r-pcc_b200/build/ab/librpcc_$REV.so"
rm -rf "$TMP"
echo "$ROOT/r-pcc_b200/build/ab/librpcc_$REV.so"
