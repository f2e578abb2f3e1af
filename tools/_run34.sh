timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "specializ" 2>&1 | tail -4
timeout 600 python tools/bench_configs.py spec 2>&1 | tail -6 | cut -c1-700
timeout 600 python tools/bench_configs.py mixed_ve mixed_ve_spec 2>&1 | tail -4 | cut -c1-400
