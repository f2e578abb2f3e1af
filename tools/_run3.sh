set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "staged_gemm" 2>&1 | tail -15
for cfg in "diabetes 2048" "munin 256" "pathfinder 16384"; do
  set -- $cfg
  timeout 600 python tools/launch_profile.py $1 $2 14 > gpurun_out/lp3_$1.txt 2>&1; head -18 gpurun_out/lp3_$1.txt
done
