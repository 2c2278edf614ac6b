#!/bin/bash
# Round-end evidence in one gpurun call: GPU test-suite, ncu launch list of the bench command, ncu --set full of the encode and
# decode kernels (profiles/README.md lists what each file shows).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; tail -3 gpurun_out/pytest_final.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02d_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pipelined > gpurun_out/r02d_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"autoc|fused|scan_kernel|compact|finalize" -c 8 -f -o gpurun_out/prof_r2d_enc \
    python tools/prof_encode.py 5 256 > gpurun_out/prof_r2d_enc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"dec_" -c 16 -f -o gpurun_out/prof_r2f_dec \
    python tools/prof_decode.py 4096 131072 > gpurun_out/prof_r2f_dec.log 2>&1
tail -n 2 gpurun_out/prof_r2d_enc.log; tail -n 2 gpurun_out/prof_r2f_dec.log
