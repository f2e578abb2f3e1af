for dbg in 0 2; do
  echo "== f32 PGX_STAGE_DEBUG=$dbg"
  PGX_STAGE_DEBUG=$dbg timeout 600 python tools/launch_profile.py diabetes 2048 6 f32 2>&1 | sed -n 2,9p
done
