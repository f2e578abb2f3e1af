timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?; tail -c 300 gpurun_out/bench_n1.err
for cfg in "diabetes 2048" "munin 256" "pathfinder 16384"; do
  set -- $cfg
  timeout 300 python tools/launch_profile.py $1 $2 40 > gpurun_out/lp7_$1.txt 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts.sum --clock-control none --csv --log-file gpurun_out/ncu7_$1.csv python tools/launch_profile.py $1 $2 40 ncu > gpurun_out/ncu7_$1.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_bench.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_plan_fused2 --launch-skip 3 --launch-count 1 -o /tmp/r02_alarm_fused -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-e2e > gpurun_out/ncu_fused.log 2>&1; echo rc=$?
python tools/ncu_raw.py /tmp/r02_alarm_fused.ncu-rep > gpurun_out/r02_alarm_fused_ncu_raw.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_contract_mm --launch-skip 2 --launch-count 2 -o /tmp/r02_mm_diabetes -f python tools/launch_profile.py diabetes 2048 8 ncu > gpurun_out/ncu_mm.log 2>&1; echo rc=$?
python tools/ncu_raw.py /tmp/r02_mm_diabetes.ncu-rep > gpurun_out/r02_mm_diabetes_ncu_raw.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_contract_tile32 --launch-skip 24 --launch-count 4 -o /tmp/r02_tile32_munin -f python tools/launch_profile.py munin 256 8 ncu > gpurun_out/ncu_tile.log 2>&1; echo rc=$?
python tools/ncu_raw.py /tmp/r02_tile32_munin.ncu-rep > gpurun_out/r02_tile32_munin_after_ncu_raw.txt
du -sh gpurun_out
