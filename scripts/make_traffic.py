"""Turn an ncu capture of scripts/profile_case.py into profiles/r02_traffic.json (run here, on the CPU box):

    python scripts/make_traffic.py gpurun_out/prof.ncu-rep gpurun_out/prof_stamp.json

The JSON holds, per decoder kernel, DRAM bytes and executed warp instructions PER FRAME (ncu `dram__bytes_read.sum +
dram__bytes_write.sum` and `smsp__inst_executed.sum` divided by the frames of the launch) and is stamped with the hash
of the kernel sources the capture ran (short_ldpc_decoding_osd_b200.build._stamp(), written next to the report by
profile_case.py on the GPU box).  bench.py refuses the file when the stamp differs from the library it is timing.
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, stamp_file = sys.argv[1], sys.argv[2]
meta = json.load(open(stamp_file))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def num(r, name, scale_unit=True):
    v = float(r[ix[name]].replace(",", ""))
    u = units[ix[name]].lower()
    if scale_unit:
        v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    return v


out = {"source": f"ncu --set full --clock-control none on scripts/profile_case.py ({meta['frames']} frames per launch), report {os.path.basename(rep)}; "
                 "dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum divided by the frames of the launch",
       "build_stamp": meta["build_stamp"], "frames": meta["frames"]}
keys = {"nms": "nms_kernel", "osd_pair": "osd_kernel", "osd_kernel": "osd_kernel_order1", "osd3": "osd_kernel_order3"}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    key = next((v for k, v in keys.items() if k in name), None)
    if key is None or key in out:
        continue
    frames = meta["frames_of"].get(key, meta["frames"])
    rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
    out[key] = {"kernel": name[:100], "frames": frames, "dram_bytes_per_frame": round((rd + wr) / frames, 1), "read_MB": round(rd / 1e6, 3),
                "write_MB": round(wr / 1e6, 3), "warp_instr_per_frame": round(num(r, "smsp__inst_executed.sum", False) / frames, 1),
                "time_us_under_ncu": round(num(r, "gpu__time_duration.sum", False) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[ix["gpu__time_duration.sum"]], 1), 1)}
dst = os.path.join(ROOT, "profiles", "r02_traffic.json")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
