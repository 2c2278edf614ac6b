#!/bin/bash
# Host logic of the drop-in decoder layer against libFLAC WITHOUT a GPU (development aid, not part of the product or of pytest).
# The shipped library refuses to initialise a decoder without a CUDA device (no CPU fallback).  To compare the parts of the handle
# API that never reach a kernel -- decoder: metadata parsing, one-process_single-per-block bookkeeping, states and return values up
# to the first audio frame; encoder: streams without a single sample (header, STREAMINFO rewrite, failing callbacks) -- with the
# reference binary on a machine without a GPU, this script builds a SCRATCH copy of the library under
# /tmp in which the two init checks (decoder, encoder) are compiled out, and runs tools/host_logic_check.py against it and oracle/_ref.  Nothing it
# builds is shipped or loaded by the package.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=${1:-/tmp/flacb200_hostcheck}
rm -rf "$W" && mkdir -p "$W/pyflac_b200"
cp -r "$ROOT/pyflac_b200/csrc" "$W/pyflac_b200/csrc" && rm -rf "$W/pyflac_b200/csrc/_obj"
cp -r "$ROOT/include" "$W/include"
cd "$W/pyflac_b200/csrc"
sed -i 's/if (!dec_ctx()) { m->state = DS_MEMORY_ALLOCATION_ERROR; return DI_MEMORY_ALLOCATION_ERROR; }/if (false) { return 0; }/' flac_api_dec.cu
grep -q 'if (false) { return 0; }' flac_api_dec.cu || { echo "init check not found in flac_api_dec.cu"; exit 1; }
sed -i 's/if (!d->context()) { m->state = ST_MEMORY_ALLOCATION_ERROR; return INIT_ENCODER_ERROR; }/if (false) { return 0; }/' flac_api_enc.cu
grep -q 'if (false) { return 0; }' flac_api_enc.cu || { echo "init check not found in flac_api_enc.cu"; exit 1; }
nvcc -gencode arch=compute_100a,code=sm_100a -O1 -std=c++17 --fmad=false -Xcompiler -fPIC -shared -cudart static \
     -o "$W/libhostcheck.so" flac_api_dec.cu dec_engine.cu dec_kernels.cu engine.cu enc_analyze.cu enc_pack.cu enc_fused.cu flac_api_enc.cu -lpthread
cd "$ROOT" && python tools/host_logic_check.py "$W/libhostcheck.so"
