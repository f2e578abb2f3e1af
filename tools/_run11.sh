timeout 120 python tools/mm_smoke.py 2>&1 | tail -12
