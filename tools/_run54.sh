timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['d2h_ceiling_value'], d['roofline']['frac'], d['gpu_launches'], d['wall_ms_per_step'])
for k,v in d['configs'].items(): print(k, v['ms_per_batch'], v['frac_of_hbm_peak'])
"
timeout 300 python tools/bench_configs.py mixed_ve_spec 2>&1 | tail -2 | cut -c1-330
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
