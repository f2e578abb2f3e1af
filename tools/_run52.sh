timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "specializ" 2>&1 | tail -2
timeout 120 python bench.py --steps 300 --warmup 20 --no-configs --no-cpu-baseline 2>gpurun_out/b52.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['ms_per_step'],4), '%.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], d['roofline']['frac'], d['engine'])"
tail -2 gpurun_out/b52.err
