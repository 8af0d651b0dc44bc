#!/bin/bash
# Builds the CPU restatement oracle (test infrastructure). -ffp-contract=off: every a*b+c is two roundings like the reference build.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
mkdir -p "$HERE/_build"
gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared -Wall -Wno-unknown-pragmas -o "$HERE/_build/libterrain_oracle.so" "$HERE/terrain_oracle.c" -lm
echo "built $HERE/_build/libterrain_oracle.so"
