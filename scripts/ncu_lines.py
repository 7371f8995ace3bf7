"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[2]
iL, iS, iA, iN, iI = 0, 1, 2, hdr.index('# Samples'), hdr.index('Instructions Executed')
iExc = hdr.index('L1 Wavefronts Shared Excessive'); iWf = hdr.index('L1 Wavefronts Shared')
lines = {}
for r in rows[3:]:
    if len(r) != len(hdr) or r[iA] != '-': continue   # per-CUDA-line aggregate rows have '-' as address
    try: lines[int(r[iL])] = (r[iS], int(r[iN]), int(r[iI]), int(r[iWf] or 0), int(r[iExc] or 0))
    except ValueError: pass
tot_i = sum(v[2] for v in lines.values()); tot_s = sum(v[1] for v in lines.values())
print('total warp instr', tot_i, 'samples', tot_s)
for ln, (src, s, i, wf, exc) in sorted(lines.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f'{ln:5d} instr {100*i/tot_i:5.1f}%  samples {100*s/max(tot_s,1):5.1f}%  smemwf {wf:9d} exc {exc:8d}  {src.strip()[:95]}')
