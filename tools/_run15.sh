for cfg in "diabetes 2048" "munin 256" "pathfinder 16384"; do
  set -- $cfg
  timeout 300 python tools/launch_profile.py $1 $2 40 > gpurun_out/lp6_$1.txt 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts.sum --clock-control none --csv --log-file gpurun_out/ncu6_$1.csv python tools/launch_profile.py $1 $2 40 ncu > gpurun_out/ncu6_$1.log 2>&1
done
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_plan_fused2 --launch-skip 3 --launch-count 1 -o gpurun_out/r02_alarm_fused -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-e2e > gpurun_out/ncu_fused.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_bench_launches.csv
