"""Per-source-line warp-stall samples from an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; lines = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0]: continue
    d = dict(zip(hdr, r))
    try: s = int(d["# Samples"])
    except Exception: continue
    # hdr has two "Source" columns; r[1] is the CUDA line text
    lines.append((s, int(r[0]), r[1].strip(), d))
tot = sum(s for s, *_ in lines) or 1
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"total samples {tot}")
for s, ln, txt, d in sorted(lines, key=lambda x: x[1]):
    if s * 200 < tot: continue   # >= 0.5 %
    st = sorted(((int(d[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{ln:5d} {100*s/tot:5.1f}%  {txt[:90]:90s} {' '.join(f'{n}:{v}' for v, n in st if v)}")
