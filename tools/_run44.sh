run() { timeout 120 python bench.py --steps 50 --warmup 10 --no-configs --no-cpu-baseline --no-e2e 2>gpurun_out/b44.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],4), '%.3e' % d['value'], d['engine']['specialized_kernel']['registers'])"; }
export PGX_SPEC_WARPS=2 PGX_SPEC_SPLIT=1 PGX_SPEC_DEBUG_SKIP_OUTPUT=1
for c in 6 7 8 10; do PGX_SPEC_MINCTAS=$c run "split compute only minctas $c"; done
