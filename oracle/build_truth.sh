#!/bin/sh
# TEST INFRASTRUCTURE: builds the extended-precision dense ground truth (oracle/dense_truth.c) twice,
# x87 long double (libmra_truth_l.so) and __float128 (libmra_truth_q.so), into oracle/_ref/ (git-ignored).
set -e
here="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$here/_ref"
gcc -O2 -fopenmp -shared -fPIC -o "$here/_ref/libmra_truth_l.so" "$here/dense_truth.c" -lm
gcc -O2 -fopenmp -shared -fPIC -DUSE_QUAD -o "$here/_ref/libmra_truth_q.so" "$here/dense_truth.c" -lquadmath -lm
