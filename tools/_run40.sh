run() { timeout 120 python bench.py --steps 50 --warmup 10 --no-configs --no-cpu-baseline --no-e2e 2>gpurun_out/b40.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],4), '%.3e' % d['value'], d['engine']['specialized_kernel'])"; }
run "stage default"
PGX_SPEC_STAGE=0 run "stage off"
for c in 4 6 7; do PGX_SPEC_MINCTAS=$c run "stage minctas $c"; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "specializ" 2>&1 | tail -3
timeout 600 python tools/bench_configs.py spec 2>&1 | tail -4 | cut -c1-330
