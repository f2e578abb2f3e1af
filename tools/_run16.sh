timeout 200 python tools/launch_profile.py diabetes 2048 3 2>&1 | head -3
timeout 200 python tools/launch_profile.py munin 256 3 2>&1 | head -3
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_abi.py -m gpu -x -q -k "matrix_product" 2>&1 | tail -3
timeout 120 python tools/mm_smoke.py | tail -1
