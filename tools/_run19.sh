timeout 300 python bench.py --steps 100 --warmup 10 --no-configs --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'])"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "jt_all_marginals or shared_memory or bp_marginals or fp32 or known_answers or ve_" 2>&1 | tail -3
timeout 300 python tools/bench_configs.py mixed_ve 2>&1 | tail -4
