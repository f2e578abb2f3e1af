for dbg in 0 1 2 3; do
  echo "== PGX_STAGE_DEBUG=$dbg"
  PGX_STAGE_DEBUG=$dbg timeout 600 python tools/launch_profile.py diabetes 2048 8 2>&1 | sed -n 2,11p
done
