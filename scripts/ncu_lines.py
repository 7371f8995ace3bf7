"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line (file:line)."""
import csv, os, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
only = sys.argv[3] if len(sys.argv) > 3 else None   # optional file-name filter
lines, hdr, fname = {}, None, '?'
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = os.path.basename(r[1]); continue
    if r[0] == 'Line No':
        hdr = r; iN, iI = hdr.index('# Samples'), hdr.index('Instructions Executed')
        iExc = hdr.index('L1 Wavefronts Shared Excessive'); iWf = hdr.index('L1 Wavefronts Shared'); continue
    if hdr is None or len(r) != len(hdr) or r[2] != '-': continue   # per-CUDA-line aggregate rows have '-' as address
    try: lines[(fname, int(r[0]))] = (r[1], int(r[iN]), int(r[iI]), int(r[iWf] or 0), int(r[iExc] or 0))
    except ValueError: pass
tot_i = sum(v[2] for v in lines.values()); tot_s = sum(v[1] for v in lines.values())
print('total warp instr', tot_i, 'samples', tot_s)
sel = {k: v for k, v in lines.items() if only is None or only in k[0]}
for (fn, ln), (src, s, i, wf, exc) in sorted(sel.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f'{fn[:16]:16s}{ln:5d} instr {100*i/tot_i:5.1f}%  samples {100*s/max(tot_s,1):5.1f}%  smemwf {wf:9d} exc {exc:8d}  {src.strip()[:90]}')
