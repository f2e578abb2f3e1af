"""Join an ncu per-launch metric list (tools/launch_profile.py ... ncu) with the plain per-launch timing table."""
import csv, json, collections, sys
name, B = sys.argv[1], int(sys.argv[2])
tag = sys.argv[3] if len(sys.argv) > 3 else "ncu4"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
rows = list(csv.reader(l for l in open(f"gpurun_out/{tag}_{name}.csv") if l.startswith('"')))
hdr, data = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "": 1, "usecond": 1e3, "nsecond": 1, "msecond": 1e6}
byid = collections.OrderedDict()
for r in data:
    d = byid.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]})
    d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * scale[r[ix["Metric Unit"]]]
ks = list(byid.values())
lp = json.load(open(f"gpurun_out/launch_profile_{name}_{B}.json"))
n = len(lp)
# last pass: n launches + emit
ks = [k for k in ks if "k_contract" in k["name"] or "k_emit" in k["name"]]
per = n + 1
last = ks[-per:]
tot_t = sum(k["gpu__time_duration.sum"] for k in last) / 1e6
tot_d = sum(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"] for k in last)
tot_a = sum(r["alg_bytes"] for r in lp)
print(f"{name} B={B}: plain {sum(r['ms'] for r in lp):.2f} ms, ncu {tot_t:.2f} ms, DRAM {tot_d/1e9:.1f} GB, alg {tot_a/1e9:.1f} GB, L2 {sum(k['lts__t_bytes.sum'] for k in last)/1e9:.1f} GB")
print("  i  kernel   ms(plain) ms(ncu)  alg GB  dramR  dramW  L2 GB  wavefrM  GB/s(alg) GB/s(dram)  desc")
out = []
for i, (r, k) in enumerate(zip(lp, last)):
    kn = "mm" if "k_contract_mm" in k["name"] else "stage" if "stage" in k["name"] else ("tile32" if "tile32" in k["name"] else ("step" if "k_contract_step" in k["name"] else k["name"][:20]))
    out.append((r["ms"], i, kn, k, r))
for ms, i, kn, k, r in sorted(out, key=lambda t: -t[0])[:top]:
    dr, dw = k["dram__bytes_read.sum"] / 1e9, k["dram__bytes_write.sum"] / 1e9
    print(f"  {i:3d} {kn:7s} {ms:8.3f} {k['gpu__time_duration.sum']/1e6:8.3f} {r['alg_bytes']/1e9:7.3f} {dr:6.3f} {dw:6.3f} {k['lts__t_bytes.sum']/1e9:6.2f} {k['l1tex__data_pipe_lsu_wavefronts.sum']/1e6:8.1f} {r['alg_bytes']/ms/1e6:8.0f} {(dr+dw)*1e3/ms:8.0f}   {r['desc']}")
# by kernel
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for ms, i, kn, k, r in out:
    a = agg[kn]; a[0] += ms; a[1] += r["alg_bytes"]; a[2] += k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"]; a[3] += 1
for kn, a in agg.items():
    print(f"  {kn}: {a[3]} launches {a[0]:.2f} ms, alg {a[1]/1e9:.1f} GB ({a[1]/a[0]/1e6:.0f} GB/s), dram {a[2]/1e9:.1f} GB ({a[2]/a[0]/1e6:.0f} GB/s)")
