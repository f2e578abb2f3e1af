timeout 120 python tools/mm_smoke.py | tail -1
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matrix_product or large_models or bp_marginals" 2>&1 | tail -3
for cfg in "diabetes 2048" "munin 256" "pathfinder 16384"; do
  set -- $cfg
  timeout 200 python tools/launch_profile.py $1 $2 5 2>&1 | head -3 | tail -2
done
