timeout 120 python tools/mm_smoke.py 2>&1 | tail -9
for d in 0 2; do
echo "== PGX_MM_DEBUG=$d"
PGX_MM_DEBUG=$d timeout 200 python tools/launch_profile.py diabetes 2048 8 2>&1 | head -11
done
timeout 200 python tools/launch_profile.py munin 256 12 2>&1 | head -15
