timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_contract_mm --launch-skip 1 --launch-count 3 -o gpurun_out/mm_diab2 -f python tools/launch_profile.py diabetes 2048 8 ncu > gpurun_out/ncu_mm.log 2>&1
tail -2 gpurun_out/ncu_mm.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_contract_tile32 --launch-skip 24 --launch-count 4 -o gpurun_out/tile_munin -f python tools/launch_profile.py munin 256 8 ncu > gpurun_out/ncu_tile.log 2>&1
tail -2 gpurun_out/ncu_tile.log
