#!/bin/bash
# One gpurun call: GPU test-suite on the library in place, then decode + encode kernel timings of the library in place and of
# every variant under _variants/.  Output: gpurun_out/ab_mixed.log
mkdir -p gpurun_out
{
  timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
  echo "== in place"; timeout 120 python tools/prof_decode.py 4096 131072; timeout 120 python tools/prof_encode.py 5 256
  timeout 500 python tools/ab_variants.py run -- bash -c 'timeout 120 python tools/prof_decode.py 4096 131072; timeout 120 python tools/prof_encode.py 5 256'
} > gpurun_out/ab_mixed.log 2>&1
cat gpurun_out/ab_mixed.log | tail -40
