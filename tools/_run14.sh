timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mpe or beliefs_vs_reference or known_answers or map_query" 2>&1 | tail -15
