timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_contract_mm --launch-skip 2 --launch-count 2 -o gpurun_out/mm_diab -f python tools/launch_profile.py diabetes 2048 8 ncu > gpurun_out/ncu_mm.log 2>&1
tail -3 gpurun_out/ncu_mm.log
ls -la gpurun_out/mm_diab*
