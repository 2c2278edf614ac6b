#!/bin/bash
# One gpurun call: GPU test-suite on the library in place, then the decode tests + decode timings of every variant under
# _variants/ (tools/ab_variants.py).  Output: gpurun_out/ab_decode.log
mkdir -p gpurun_out
{
  timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
  timeout 500 python tools/ab_variants.py run -- bash -c 'timeout 120 python -m pytest tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -2; timeout 120 python tools/prof_decode.py 4096 131072; timeout 120 python tools/prof_decode.py 256 480000'
} > gpurun_out/ab_decode.log 2>&1
tail -30 gpurun_out/ab_decode.log
