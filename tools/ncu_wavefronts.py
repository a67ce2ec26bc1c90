"""Per-source-line shared-memory wavefronts / executed instructions from an ncu report (-lineinfo + --import-source on).
usage: python tools/ncu_wavefronts.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; lines = []; fpath = ""
for r in rows:
    if r and r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0]: continue
    d = dict(zip(hdr, r))
    try:
        lines.append((int(d["L1 Wavefronts Shared"] or 0), int(d["L1 Wavefronts Shared Ideal"] or 0), int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0), fpath, int(r[0]), r[1].strip()))
    except Exception: continue
tw = sum(l[0] for l in lines) or 1; ti = sum(l[2] for l in lines) or 1; ts = sum(l[3] for l in lines) or 1
print(f"shared wavefronts {tw} (ideal {sum(l[1] for l in lines)}), warp instructions {ti}, samples {ts}")
print("  wavefronts   %   ideal | instr %  | samples % | line")
for w, wi, ins, s, f, ln, txt in sorted(lines, key=lambda x: -x[0])[:top]:
    print(f"{w:10d} {100*w/tw:5.1f} {wi:9d} | {100*ins/ti:5.1f} | {100*s/ts:5.1f} | {f}:{ln} {txt[:80]}")
