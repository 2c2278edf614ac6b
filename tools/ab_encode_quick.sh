#!/bin/bash
# One gpurun call: encode parity tests on the library in place, then encode kernel timings in place and of every variant.
mkdir -p gpurun_out
{
  timeout 300 python -m pytest tests/test_gpu_encode.py tests/test_gpu_fixtures.py -m gpu -x -q 2>&1 | tail -4
  echo "== in place"; timeout 120 python tools/prof_encode.py 5 256
  timeout 300 python tools/ab_variants.py run -- bash -c 'timeout 120 python tools/prof_encode.py 5 256'
} > gpurun_out/ab_encode_quick.log 2>&1
cat gpurun_out/ab_encode_quick.log
