#!/bin/bash
# One gpurun call: GPU test-suite on the library in place, then decode timings of every variant under _variants/
# (tools/ab_variants.py; base = the previous build).  Output: gpurun_out/ab_decode.log
mkdir -p gpurun_out
{
  timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
  timeout 300 python tools/ab_variants.py run -- bash -c 'timeout 120 python tools/prof_decode.py 4096 131072; timeout 120 python tools/prof_decode.py 256 480000'
} > gpurun_out/ab_decode.log 2>&1
tail -30 gpurun_out/ab_decode.log
