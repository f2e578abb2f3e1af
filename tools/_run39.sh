run() { timeout 120 python bench.py --steps 50 --warmup 10 --no-configs --no-cpu-baseline --no-e2e 2>gpurun_out/b39.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],4), '%.3e' % d['value'])"; }
run "default"
PGX_SPEC_DEBUG_SKIP_COMPUTE=1 run "output stage only"
PGX_SPEC_DEBUG_SKIP_OUTPUT=1 run "compute only"
python - <<'PY'
import torch,time
x=torch.empty(131072*91,dtype=torch.float64,device='cuda')
for _ in range(5): x.fill_(1.0)
torch.cuda.synchronize()
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50): x.fill_(1.0)
b.record(); torch.cuda.synchronize()
print('torch fill 95MB ms', a.elapsed_time(b)/50)
PY
