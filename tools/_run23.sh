timeout 200 python tools/launch_profile.py diabetes 2048 6 2>&1 | head -9
echo "== with a stream sync before every launch"
PGX_PROFILE_SYNC=1 timeout 200 python tools/launch_profile.py diabetes 2048 6 2>&1 | head -9
