"""Trim an `ncu --metrics gpu__time_duration.sum --csv` launch list into profiles/ form + per-kernel shares.
usage: python tools/launches_summary.py gpurun_out/launches.csv "<command that was profiled>" ["(1024, 1, 1)"] > profiles/xxx.csv
The optional third argument restricts the share computation to launches with that grid (the headline step's launches:
bench.py also runs the host-buffer e2e leg, whose chunks launch the same kernels on smaller grids)."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
print(f"# {sys.argv[2] if len(sys.argv) > 2 else ''} (ns per launch; cold-cache, serialised: compare shares, not absolutes)")
print("id,kernel,block,grid,duration_ns")
agg = collections.OrderedDict()
gridsel = sys.argv[3] if len(sys.argv) > 3 else None
sel = collections.OrderedDict()
for r in rows:
    k = r["Kernel Name"]
    short = k.split("(")[0].replace("<unnamed>::", "").replace("void ", "")[:60]
    print(f'{r["ID"]},"{short}","{r["Block Size"]}","{r["Grid Size"]}",{r["Metric Value"]}')
    agg.setdefault(short, []).append(float(r["Metric Value"]))
    if gridsel and r["Grid Size"].replace(" ", "") == gridsel.replace(" ", ""):
        sel.setdefault(short, []).append(float(r["Metric Value"]))
step = {k: v for k, v in (sel if gridsel else agg).items() if k.startswith("k_") and not k.startswith("k_peak")}
tot = sum(sum(v) for v in step.values()) or 1.0
print("# per-kernel totals over the captured launches; share = share of the solve step (this library's k_* kernels, microbenchmarks excluded)")
if gridsel:
    print(f"# shares over the launches with grid {gridsel} only (the headline step)")
    for k, v in step.items():
        print(f"# {k:60s} launches {len(v):4d}  mean {sum(v)/len(v):12.0f} ns  share {100*sum(v)/tot:5.1f} %")
    print("# all launches:")
for k, v in agg.items():
    sh = (f"{100*sum(v)/tot:5.1f} %" if k in step else "  (not part of the step)") if not gridsel else ""
    print(f"# {k:60s} launches {len(v):4d}  mean {sum(v)/len(v):12.0f} ns  share {sh}")
