timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['d2h_ceiling_value'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'])
for k,v in d['configs'].items(): print(k, v['ms_per_batch'], v['frac_of_hbm_peak'])
print(d['cpu_baseline'])
"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
