python -m pytest tests/test_gpu_decode.py tests/test_gpu_fixtures.py tests/test_gpu_dropin.py tests/test_gpu_pyapi.py -m gpu -x -q 2>&1 | tail -3
python tools/prof_decode.py 128 2>&1 | tail -1
python tools/prof_decode.py 4096 131072 2>&1 | tail -1
