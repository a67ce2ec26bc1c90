"""Executed instructions and stall samples per source line of one file, from an ncu report (needs -lineinfo, --import-source on).
usage: python tools/ncu_inst.py report.ncu-rep file_substring [min_pct]"""
import csv, subprocess, sys
rep, want = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; agg = {}
for r in rows:
    if r and r[0] in ("File Name", "File Path"):
        cur = r[1]; hdr = None; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or cur is None or want not in cur or len(r) < len(hdr) or not r[0]: continue
    d = dict(zip(hdr, r))
    try:
        ln = int(r[0]); ins = int(d["Instructions Executed"] or 0); smp = int(d["# Samples"] or 0)
    except Exception:
        continue
    a = agg.setdefault(ln, [0, 0, r[1].strip(), {}])
    a[0] += ins; a[1] += smp
    for k in d:
        if k.startswith("stall_") and "Not Issued" not in k and d[k]:
            try: a[3][k[6:]] = a[3].get(k[6:], 0) + int(d[k])
            except Exception: pass
ti = sum(a[0] for a in agg.values()) or 1; ts = sum(a[1] for a in agg.values()) or 1
print(f"file {want}: instructions executed {ti}, samples {ts}")
for ln in sorted(agg):
    i, s, txt, st = agg[ln]
    if 100 * i / ti < minpct and 100 * s / ts < minpct: continue
    top = " ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{ln:5d} inst {100*i/ti:5.1f}%  smp {100*s/ts:5.1f}%  {txt[:80]:80s} {top}")
