run() { timeout 120 python bench.py --steps 50 --warmup 10 --no-configs --no-cpu-baseline --no-e2e 2>gpurun_out/b38.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],4), '%.3e' % d['value'], d['engine']['specialized_kernel'])"; }
run "default R=1"
for r in 2 4 8; do for sy in 2 6 20; do PGX_SPEC_ROWS=$r PGX_SPEC_SYNC=$sy run "rows $r sync $sy"; done; done
PGX_SPEC_ROWS=8 PGX_SPEC_SYNC=100000 run "rows 8 no sync"
PGX_SPEC_ROWS=8 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "specializ" 2>&1 | tail -3
