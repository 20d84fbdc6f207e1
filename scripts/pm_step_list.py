"""Per-launch list of the LAST PredictiveModel step in an ncu launch list (gpurun_out/pm_launches.csv from scripts/gpu_pm_iter.sh),
torch's own fill kernels left out; with --agg the per-kernel totals."""
import csv, re, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr, data = rows[0], rows[1:]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
idx = [i for i, r in enumerate(data) if "pm_loss_final" in r[ki]]
L = idx[-1] - idx[-2]
step = data[len(data) - L:]
tot, agg = 0.0, collections.OrderedDict()
for i, r in enumerate(step):
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "").replace("avc::", "")
    if name.startswith("at::"): continue
    t = float(r[vi].replace(",", "")) / 1e3; tot += t
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
    if "--agg" not in sys.argv: print(f"{i:3d} {name:28s} {r[gi]:>16s} {t:8.1f}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(f"{k:28s} {n:3d} {t:8.1f} us {100*t/tot:5.1f}%")
print("total", round(tot, 1), "us")
