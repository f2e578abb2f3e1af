timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matrix_product_tile" 2>&1 | tail -8
for cfg in "diabetes 2048" "munin 256" "pathfinder 16384"; do
  set -- $cfg
  timeout 200 python tools/launch_profile.py $1 $2 16 > gpurun_out/lp5_$1.txt 2>&1
  head -14 gpurun_out/lp5_$1.txt
done
