#!/bin/bash
# md5 of the SASS text of every kernel in rdr_kernels.cu (opcodes and operands, no encodings, no line info): two builds
# with the same fingerprint run the same device code.  Used to show that prepared, default-off variants (RDR_CHUNKED,
# RDR_DIRECT_BALLOT, RDR_APPROX_RHO) leave the validated default build untouched.
# Usage: bash scripts/sass_fingerprint.sh [extra nvcc flags, e.g. -DRDR_DIRECT_BALLOT=1]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp /tmp/rdr_fp_XXXX.cubin)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 "$@" -I "$ROOT/include" -I "$ROOT/raydar_b200/csrc" -cubin -o "$TMP" "$ROOT/raydar_b200/csrc/rdr_kernels.cu"
cuobjdump -sass "$TMP" | grep -E "^\s+/\*[0-9a-f]{4}\*/|Function" | sed 's/ *\/\* 0x[0-9a-f]* \*\///' | md5sum | cut -d' ' -f1
rm -f "$TMP"
