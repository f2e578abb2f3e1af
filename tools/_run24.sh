echo "== level streams off"; PGX_LEVEL_STREAMS=0 timeout 300 python tools/time_configs.py
echo "== level streams on"; timeout 300 python tools/time_configs.py
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_models or matrix_product or graph or bp_marginals or pinned" 2>&1 | tail -3
