python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['d2h_ceiling_value'], d['engine']['kernel_variant'], d.get('posterior_gather'))
"
