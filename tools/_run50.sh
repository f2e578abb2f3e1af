timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err
ncu --set full --clock-control none --import-source on -k regex:k_plan_spec -c 1 -s 3 -o gpurun_out/r02_alarm_spec -f python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/ncu_spec.log 2>&1
python tools/ncu_raw.py gpurun_out/r02_alarm_spec.ncu-rep > gpurun_out/r02_alarm_spec_ncu_raw.txt 2>&1
cat gpurun_out/r02_alarm_spec_ncu_raw.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
grep -c k_plan_spec gpurun_out/r02_bench_launches.csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/bench_configs.py spec mixed_ve_spec 2>&1 | tail -6 | cut -c1-420
