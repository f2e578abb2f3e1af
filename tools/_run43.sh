run() { timeout 120 python bench.py --steps 50 --warmup 10 --no-configs --no-cpu-baseline --no-e2e 2>gpurun_out/b43.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],4), '%.3e' % d['value'], d['engine']['specialized_kernel']['registers'])"; }
for a in 1 2 3 4; do PGX_SPEC_ACC=$a run "acc $a"; done
for a in 2 3; do PGX_SPEC_ACC=$a PGX_SPEC_DEBUG_SKIP_OUTPUT=1 run "acc $a compute only"; done
