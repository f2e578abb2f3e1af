for cfg in "diabetes 2048" "munin 256" "pathfinder 16384"; do
  set -- $cfg
  python tools/launch_profile.py $1 $2 40 > gpurun_out/lp4_$1.txt 2>&1
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts.sum --clock-control none --csv --log-file gpurun_out/ncu4_$1.csv python tools/launch_profile.py $1 $2 40 ncu > gpurun_out/ncu4_$1.log 2>&1
done
tail -3 gpurun_out/ncu4_*.log
