#!/bin/bash
# One gpurun call: decode timings (4096 x 131072 and 256 x 480000) of every variant under _variants/.  Output: gpurun_out/ab_decode_quick.log
mkdir -p gpurun_out
timeout 500 python tools/ab_variants.py run -- bash -c 'timeout 120 python tools/prof_decode.py 4096 131072; timeout 120 python tools/prof_decode.py 256 480000' > gpurun_out/ab_decode_quick.log 2>&1
cat gpurun_out/ab_decode_quick.log
