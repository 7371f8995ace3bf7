"""Text report of one kernel of an .ncu-rep (run here, on the CPU box):

    python scripts/ncu_kernel_report.py gpurun_out/prof.ncu-rep <kernel regex> <frames per launch> > profiles/r02_ncu_<name>.txt

Key metrics, stall reasons (warp-state samples), executed opcode mix, and the hottest CUDA source lines."""
import collections, csv, io, re, subprocess, sys
rep, kre, frames = sys.argv[1], sys.argv[2], float(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "-k", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
ix = {h: i for i, h in enumerate(hdr)}
def val(m):
    try: return float(r[ix[m]].replace(",", ""))
    except Exception: return None
print(f"## {r[ix['Kernel Name']]}  grid {r[ix['launch__grid_size']]} block {r[ix['launch__block_size']]}  ({int(frames)} frames in this launch)")
M = [("gpu__time_duration.sum", "time under ncu"), ("smsp__inst_executed.sum", "warp instructions"), ("sm__inst_executed.avg.per_cycle_elapsed", "IPC per SM (of 4)"),
     ("smsp__issue_active.avg.pct", "issue slots busy %"), ("sm__warps_active.avg.per_cycle_active", "warps active per SM"), ("launch__registers_per_thread", "registers"),
     ("launch__occupancy_limit_registers", "CTAs/SM by registers"), ("launch__occupancy_limit_shared_mem", "CTAs/SM by shared memory"),
     ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
     ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts (incl. shuffles)"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
     ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"), ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"), ("sm__inst_executed_pipe_alu.sum", "ALU-pipe instructions"),
     ("sm__inst_executed_pipe_fma.sum", "FMA-pipe instructions"), ("sm__inst_executed_pipe_lsu.sum", "LSU-pipe instructions")]
for m, label in M:
    if m in ix and r[ix[m]] not in ("", "n/a"):
        extra = ""
        if m.endswith("executed.sum") or "wavefronts" in m or "pipe_alu.sum" in m or "pipe_fma.sum" in m or "pipe_lsu.sum" in m:
            v = val(m)
            if v is not None: extra = f"   ({v / frames:.0f} per frame)"
        if m.startswith("dram__bytes"):
            v = val(m)
            sc = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(units[ix[m]].lower(), 1)
            if v is not None: extra = f"   ({v * sc / frames:.1f} B per frame)"
        print(f"  {label:42s} {r[ix[m]]} {units[ix[m]]}{extra}")
st = [(val(h), h) for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and val(h) is not None]
tot = sum(v for v, _ in st) or 1
print("\nstall reasons (share of warp-state samples):")
for v, n in sorted(st, reverse=True)[:10]:
    print(f"  {100 * v / tot:5.1f}%  {n.replace('smsp__pcsamp_warps_issue_stalled_', '')}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kre], capture_output=True).stdout.decode("utf-8", "replace")
rows = list(csv.reader(io.StringIO(src)))
ops, lines, hdr2, fname = collections.Counter(), {}, None, "?"
for q in rows:
    if not q: continue
    if q[0] == "File Path": fname = q[1].split("/")[-1]; continue
    if q[0] == "Line No" or q[0] == "Address":
        hdr2 = q; continue
    if hdr2 is None or len(q) != len(hdr2): continue
    try:
        iI, iS = hdr2.index("Instructions Executed"), hdr2.index("# Samples")
    except ValueError:
        continue
    if hdr2[0] == "Line No" and q[2] == "-":
        try: lines[(fname, int(q[0]))] = (q[1], int(q[iS]), int(q[iI]))
        except ValueError: pass
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre], capture_output=True).stdout.decode("utf-8", "replace")
srows = list(csv.reader(io.StringIO(sass)))
sh = next((q for q in srows if "Source" in q and "Instructions Executed" in q), None)
if sh:
    iSrc, iIe = sh.index("Source"), sh.index("Instructions Executed")
    for q in srows:
        if len(q) == len(sh) and q[0].startswith("0x"):
            tok = [o for o in q[iSrc].split() if not o.startswith("@")]
            try: ops[tok[0].split(".")[0] if tok else "?"] += int(q[iIe])
            except ValueError: pass
ti = sum(v[2] for v in lines.values()) or 1; ts = sum(v[1] for v in lines.values()) or 1
if ops:
    to = sum(ops.values())
    print("\nexecuted opcode mix: " + ", ".join(f"{o} {100 * c / to:.1f}%" for o, c in ops.most_common(16)))
print("\nhottest CUDA lines (share of executed instructions / of samples):")
for (fn, ln), (text, s, i) in sorted(lines.items(), key=lambda kv: -kv[1][2])[:24]:
    print(f"  {fn[:18]:18s}{ln:5d}  instr {100 * i / ti:5.1f}%  samples {100 * s / ts:5.1f}%  {text.strip()[:100]}")
