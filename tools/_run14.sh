timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "soft_evidence or virtual_evidence or map_query" 2>&1 | tail -15
