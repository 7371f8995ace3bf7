"""Per-kernel SASS opcode histogram of the shipped library (runs on the CPU box: cuobjdump only).

    python scripts/sass_opcodes.py > profiles/r02_sass_opcodes.txt

Lists, for every kernel in libldpc_b200.so, the instruction count, the opcodes that matter for the Blackwell story
(legacy tensor path IMMA/HMMA, warp-reduce REDUX, 3-input min/max FMNMX3/VIMNMX3, shuffles, shared-memory traffic) and
the counts of the tcgen05 / TMEM / TMA mnemonics (UTC*MMA, LDTM, STTM, UTMALDG, UTMASTG, UBLKCP), which are zero here
by design."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "short_ldpc_decoding_osd_b200", "libldpc_b200.so")
sys.path.insert(0, ROOT)
from short_ldpc_decoding_osd_b200 import build
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
elf = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ["IMMA", "HMMA", "REDUX", "FMNMX3", "VIMNMX3", "FMNMX", "VIMNMX", "SHFL", "LDS", "STS", "LDG", "STG", "ATOMG", "RED", "BAR", "MUFU"]
BW = ["UTCHMMA", "UTCIMMA", "UTCQMMA", "UTCMXQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS"]
print(f"# SASS opcode histogram of libldpc_b200.so, build stamp {build._stamp()[:16]}")
print("# embedded cubins:", ", ".join(sorted(set(re.findall(r"sm_\d+a?", elf)))))
print("# columns: static instruction counts (not executed counts)\n")
tot = collections.Counter()
for k, c in hist.items():
    n = sum(c.values())
    tot.update(c)
    print(f"{k}\n    instructions {n}")
    print("    " + "  ".join(f"{o}:{c[o]}" for o in KEY if c[o]))
    print("    tcgen05/TMEM/TMA: " + "  ".join(f"{o}:{c[o]}" for o in BW))
    print("    top: " + "  ".join(f"{o}:{v}" for o, v in c.most_common(10)))
print("\n# whole library")
print("  " + "  ".join(f"{o}:{tot[o]}" for o in KEY))
print("  tcgen05/TMEM/TMA mnemonics: " + "  ".join(f"{o}:{tot[o]}" for o in BW), "(any UTC*:", sum(v for o, v in tot.items() if o.startswith("UTC")), ")")
