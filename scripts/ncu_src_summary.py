"""Summarise an `ncu --page source --csv` dump: stall totals and the hottest SASS lines."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) == len(hdr) and r[0].startswith('0x'):
        data.append(r)
    elif r and r[0] == 'Kernel Name':
        break  # only the first kernel of the dump
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter()
for r in data:
    for s in stalls:
        try: tot[s] += int(r[idx[s]])
        except: pass
allsamp = sum(tot.values())
print('total samples', allsamp)
for s, v in tot.most_common(10): print(f'  {s:28s} {v:8d} {100*v/allsamp:5.1f}%')
ins = sum(int(r[idx['Instructions Executed']]) for r in data)
print('instructions executed (warp-level)', ins)
opc = collections.Counter()
for r in data:
    op = r[idx['Source']].split()
    op = [o for o in op if not o.startswith('@')][0].split('.')[0] if op else '?'
    opc[op] += int(r[idx['Instructions Executed']])
print('opcode mix:', ', '.join(f'{k}:{100*v/ins:.1f}%' for k, v in opc.most_common(14)))
bank = sum(int(r[idx['L1 Wavefronts Shared Excessive']] or 0) for r in data)
ideal = sum(int(r[idx['L1 Wavefronts Shared Ideal']] or 0) for r in data)
print('shared wavefronts excessive/ideal', bank, ideal)
print('hottest lines:')
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:top]:
    st = {s: int(r[idx[s]]) for s in stalls if int(r[idx[s]] or 0)}
    main = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(f"  {int(r[idx['# Samples']]):6d} {r[idx['Source']].strip()[:70]:70s} exc_wf={r[idx['L1 Wavefronts Shared Excessive']]:>7s} {main}")
