timeout 120 python bench.py --steps 50 --warmup 10 --no-configs --no-cpu-baseline --no-e2e 2>gpurun_out/b48.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['ms_per_step'],4), '%.3e' % d['value'], d['roofline']['frac'])"
ncu --set full --clock-control none --import-source on -k regex:k_plan_spec -c 1 -s 3 -o gpurun_out/r02_alarm_spec -f python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/ncu_spec.log 2>&1
python tools/ncu_raw.py gpurun_out/r02_alarm_spec.ncu-rep > gpurun_out/r02_alarm_spec_ncu_raw.txt 2>&1
grep -E "duration|issue_active|fp64|stalls|inst_executed.sum" gpurun_out/r02_alarm_spec_ncu_raw.txt
