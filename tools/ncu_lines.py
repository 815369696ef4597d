"""Aggregate executed warp-instructions and stall samples per source line from an .ncu-rep (needs -lineinfo)."""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
cur = None
agg = collections.defaultdict(lambda: [0, 0])
files = collections.defaultdict(lambda: [0, 0])
hdr = None
started = {}
first_file = None
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == 'File Path':
        cur = row[1].split('/')[-1]
        if cur in started and started[cur] == 'done':
            cur = None          # second kernel instance of the same report: stop after the first
        continue
    if row[0] == 'Function Name':
        continue
    if row[0] == 'Line No':
        hdr = row; continue
    if hdr is None or cur is None:
        continue
    if len(row) > 2 and row[2] != '-':
        continue                # SASS row; the per-line aggregate has '-' as its address
    try:
        ln = int(row[0])
        inst = int(row[hdr.index('Instructions Executed')])
        samp = int(row[hdr.index('# Samples')])
    except (ValueError, IndexError):
        continue
    agg[(cur, ln, row[1].strip()[:90])][0] += inst
    agg[(cur, ln, row[1].strip()[:90])][1] += samp
    files[cur][0] += inst; files[cur][1] += samp
tot = sum(v[0] for v in files.values()) or 1
tots = sum(v[1] for v in files.values()) or 1
print("per file:")
for f, v in sorted(files.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:28s} inst {v[0]:>14d} {v[0]/tot:6.1%}   samples {v[1]:>8d} {v[1]/tots:6.1%}")
print("top lines by executed instructions:")
for (f, ln, src), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"  {v[0]/tot:6.2%} inst  {v[1]/tots:6.2%} stall  {f}:{ln}  {src}")
