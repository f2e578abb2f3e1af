set -x
timeout 120 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/b32.json 2>gpurun_out/b32.err
ncu --set full --clock-control none --import-source on -k regex:k_plan_spec -c 1 -s 3 -o gpurun_out/r02_alarm_spec -f python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/ncu_spec.log 2>&1
python tools/ncu_raw.py gpurun_out/r02_alarm_spec.ncu-rep > gpurun_out/r02_alarm_spec_ncu_raw.txt 2>&1
cat gpurun_out/r02_alarm_spec_ncu_raw.txt
