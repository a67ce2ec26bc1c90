"""Summarise an ncu report's source page: stall samples per SASS window (development aid)."""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]; W = int(sys.argv[2]) if len(sys.argv) > 2 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]; data = rows[2:]
iS = hdr.index('# Samples'); iSrc = hdr.index('Source'); iEx = hdr.index('Instructions Executed')
tot = sum(int(r[iS]) for r in data)
print('total samples', tot, 'ninstr', len(data))
def op(r):
    t = r[iSrc].split()
    return t[1] if t[0].startswith('@') else t[0]
for k in range(0, len(data), W):
    win = data[k:k+W]
    s = sum(int(r[iS]) for r in win)
    ops = collections.Counter(op(r) for r in win)
    ex = max(int(r[iEx]) for r in win)
    print(f'{k:5d} {100*s/tot:6.2f}%  maxexec {ex:>10d}  ' + ', '.join(f'{o}:{c}' for o,c in ops.most_common(6)))
