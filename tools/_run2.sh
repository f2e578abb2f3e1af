set -x
for cfg in "diabetes 2048" "munin 256" "pathfinder 16384"; do
  set -- $cfg
  python tools/launch_profile.py $1 $2 12 > gpurun_out/lp2_$1.txt 2>&1; head -16 gpurun_out/lp2_$1.txt
  python tools/launch_profile.py $1 $2 12 gemm > gpurun_out/lp2_$1_gemm.txt 2>&1; head -8 gpurun_out/lp2_$1_gemm.txt
done
python tools/bench_configs.py one hepar2 32768 > gpurun_out/bc_hepar2.txt 2>&1; cat gpurun_out/bc_hepar2.txt
python tools/bench_configs.py one win95pts 32768 > gpurun_out/bc_win95pts.txt 2>&1; cat gpurun_out/bc_win95pts.txt
python tools/bench_configs.py one diabetes 2048 > gpurun_out/bc_diabetes.txt 2>&1; cat gpurun_out/bc_diabetes.txt
python tools/bench_configs.py one munin 256 > gpurun_out/bc_munin.txt 2>&1; cat gpurun_out/bc_munin.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
