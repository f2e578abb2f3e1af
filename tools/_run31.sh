run() { timeout 120 python bench.py --steps 50 --warmup 10 --no-configs --no-cpu-baseline --no-e2e 2>gpurun_out/b31.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],4), '%.3e' % d['value'], d['engine']['kernel_variant'], d['engine']['specialized_kernel'])"; }
for c in 9 8 7 6 5 4 3; do PGX_SPEC_WARPS=1 PGX_SPEC_MINCTAS=$c run "warps 1 minctas $c"; done
for w in 2 4; do PGX_SPEC_WARPS=$w run "warps $w"; done
PGX_SPEC_WARPS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "specializ" 2>&1 | tail -3
