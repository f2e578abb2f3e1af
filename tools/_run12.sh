timeout 120 python tools/mm_smoke.py 2>&1 | tail -2
timeout 200 python tools/launch_profile.py diabetes 2048 4 2>&1 | head -7
timeout 200 python tools/launch_profile.py munin 256 6 2>&1 | head -9
