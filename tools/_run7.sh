python bench.py --steps 50 --warmup 10 --no-configs --no-cpu-baseline > gpurun_out/bench_fused2.json 2> gpurun_out/bench_fused2.err; echo rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/bench_fused2.json').read().strip().splitlines()[-1]); print('alarm ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "jt_all_marginals or shared_memory or bp_marginals or fp32 or known_answers" 2>&1 | tail -4
