run() { timeout 120 python bench.py --steps 50 --warmup 10 --no-configs --no-cpu-baseline --no-e2e 2>gpurun_out/b37.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],4), '%.3e' % d['value'], d['engine']['kernel_variant'], d['engine']['specialized_kernel'])"; }
run "default G=1"
export PGX_SPEC_WARPS=2 PGX_SPEC_SPLIT=1
for c in 4 5 6 7 8 10; do PGX_SPEC_MINCTAS=$c run "split minctas $c"; done
PGX_SPEC_MINCTAS=6 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "specializ" 2>&1 | tail -3
