timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?; tail -c 300 gpurun_out/bench_n1.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "graph or pinned or large_models" 2>&1 | tail -2
