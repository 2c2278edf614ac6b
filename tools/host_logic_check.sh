#!/bin/bash
# Host logic of the drop-in layer against libFLAC WITHOUT a GPU (development aid and CPU test; not part of the product).
# The shipped library refuses to initialise an encoder or decoder handle without a CUDA device (no CPU fallback).  To compare the
# parts of the handle API that never reach a kernel -- decoder: the search for the stream marker, metadata parsing and callbacks,
# one-process_single-per-block bookkeeping, getters, states and return values up to the first audio frame; encoder: streams without
# a single sample (header, STREAMINFO rewrite, failing callbacks) -- with the reference binary on a machine without a GPU, this
# script builds a SCRATCH copy of the library under /tmp in which the two init checks (decoder, encoder) are compiled out, and runs
# tools/host_logic_check.py against it and oracle/_ref.  Only the two files of the handle API are compiled (the kernels come from the
# product's own objects when they are there).  Nothing it builds is shipped or loaded by the package.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=${1:-/tmp/flacb200_hostcheck}
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
rm -rf "$W" && mkdir -p "$W/pyflac_b200"
cp -r "$ROOT/pyflac_b200/csrc" "$W/pyflac_b200/csrc" && rm -rf "$W/pyflac_b200/csrc/_obj"
cp -r "$ROOT/include" "$W/include"
cd "$W/pyflac_b200/csrc"
sed -i 's/if (!dec_ctx()) { m->state = DS_MEMORY_ALLOCATION_ERROR; return DI_MEMORY_ALLOCATION_ERROR; }/if (false) { return 0; }/' flac_api_dec.cu
grep -q 'if (false) { return 0; }' flac_api_dec.cu || { echo "init check not found in flac_api_dec.cu"; exit 1; }
sed -i 's/if (!d->context()) { m->state = ST_MEMORY_ALLOCATION_ERROR; return INIT_ENCODER_ERROR; }/if (false) { return 0; }/' flac_api_enc.cu
grep -q 'if (false) { return 0; }' flac_api_enc.cu || { echo "init check not found in flac_api_enc.cu"; exit 1; }
FLAGS="-gencode arch=compute_100a,code=sm_100a -O1 -std=c++17 --fmad=false -Xcompiler -fPIC"
OBJ="$ROOT/pyflac_b200/csrc/_obj"
REST="engine enc_analyze enc_pack enc_fused dec_kernels dec_engine"
have=1; for f in $REST; do [ -f "$OBJ/$f.cu.o" ] && [ "$OBJ/$f.cu.o" -nt "$ROOT/pyflac_b200/csrc/$f.cu" ] || have=0; done
if [ $have = 1 ]; then
    $NVCC $FLAGS -c -o flac_api_dec.o flac_api_dec.cu &
    $NVCC $FLAGS -c -o flac_api_enc.o flac_api_enc.cu
    wait
    $NVCC -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o "$W/libhostcheck.so" flac_api_dec.o flac_api_enc.o $(for f in $REST; do echo "$OBJ/$f.cu.o"; done) -lpthread
else
    $NVCC $FLAGS -shared -cudart static -o "$W/libhostcheck.so" flac_api_dec.cu flac_api_enc.cu $(for f in $REST; do echo "$f.cu"; done) -lpthread
fi
cd "$ROOT" && python tools/host_logic_check.py "$W/libhostcheck.so"
