timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for cfg in "diabetes 2048" "munin 256" "pathfinder 16384"; do
  set -- $cfg
  timeout 200 python tools/launch_profile.py $1 $2 6 2>&1 | head -9
done
