run() { timeout 120 python bench.py --steps 50 --warmup 10 --no-configs --no-cpu-baseline --no-e2e 2>gpurun_out/b42.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],4), '%.3e' % d['value'], d['engine']['specialized_kernel']['registers'])"; }
run "stage default"
PGX_SPEC_DEBUG_SKIP_OUTPUT=1 run "stage compute only"
PGX_SPEC_DEBUG_SKIP_COMPUTE=1 run "stage output only"
PGX_SPEC_DEBUG_NOWAIT=1 run "stage no wait"
