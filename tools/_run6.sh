python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?; tail -c 600 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?; tail -c 300 gpurun_out/bench_ref.err
nproc
