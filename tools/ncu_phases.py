#!/usr/bin/env python
"""Stall samples / executed instructions of an .ncu-rep source page, split at marker instructions
(barriers, MMA batches, first gather load, first reduction) so the time of each kernel phase can be read off."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]; ix = {n: i for i, n in enumerate(h)}
data = rows[hi + 1:]
def f(r, n):
    try: return float(r[ix[n]])
    except Exception: return 0.0
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
tot = sum(f(r, '# Samples') for r in data)
marks = []
prev = ''
for i, r in enumerate(data):
    s = r[ix['Source']]
    key = None
    for k in ('BAR.SYNC', 'UTCHMMA', 'LDG.E.128', 'REDG', 'SYNCS.PHASECHK', 'LDTM', 'UTCBAR'):
        if k in s: key = k
    if key and key != prev and f(r, 'Instructions Executed') > 0:
        marks.append((i, key)); prev = key
    elif key is None and False:
        prev = ''
marks.append((len(data), 'end'))
print(f"total samples {int(tot)}, SASS instructions {len(data)}")
start, name = 0, 'prologue'
ninst_ref = float(sys.argv[2]) if len(sys.argv) > 2 else max(f(r, "Instructions Executed") for r in data)
for i, key in marks:
    seg = data[start:i]
    if seg:
        smp = sum(f(r, '# Samples') for r in seg); ie = sum(f(r, 'Instructions Executed') for r in seg)
        top = sorted(stalls, key=lambda n: -sum(f(r, n) for r in seg))[:2]
        tops = ', '.join(f"{n[6:]} {sum(f(r, n) for r in seg) / max(smp, 1):.2f}" for n in top)
        print(f"[{start:5d},{i:5d}) until {key:14s} samples {int(smp):6d} {smp / tot:6.3f}  warp-instr {ie / ninst_ref:8.1f} x{int(ninst_ref)}  {tops}")
    start = i
