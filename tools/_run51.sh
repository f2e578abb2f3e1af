timeout 600 python tools/bench_configs.py spec 2>&1 | tail -4 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    f=d['flops_plan']
    print(d['config'][:22], 'default-plan spec ms', round(d['ms_specialized'],4), '| flops-plan', f['distribute'], f['factorized'], 'ms', round(f['ms'],4), f['variant'], 'err', f['max_rel_diff_vs_default'], 'regs', f['spec']['registers'], 'fp', f['spec']['fp_instr_per_row'], 's', f['plan_and_specialize_s'])
"
