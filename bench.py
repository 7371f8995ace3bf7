#!/usr/bin/env python
"""Headline benchmark: decoded frames/s of the hot path (NMS 12 iterations + order-2 OSD on the NMS
failures, CCSDS (128,64), Eb/N0 = 2.5 dB) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W             # our arm, one JSON line on rank 0
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm (C port of the reference path)

A step = one pass of the whole device pipeline (ldpcb_decode: NMS on all frames, compaction of the
detected failures, order-2 OSD on them, tallies) over one batch of B synthetic frames per GPU that is
already resident in HBM (generated once by the Philox kernel), followed by the all-reduce of the 16
uint64 counters.  The batch (B * 512 B = 1 GiB at the default B = 2^21) is larger than L2 (126 MB), so
no explicit flush is needed between steps.  `e2e` is the same pipeline through the host-buffer C-ABI
call (ldpcb_decode_host: pinned host LLRs in, decisions + counters out, copies inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALPHA = 0.66943514  # softplus(-0.048), the reference's initial NMS weight (ms_test.py:73,207-208)
WORKLOAD = "ccsds128x64_nms12_plus_osd2_on_failures_ebn0_2.5dB"
METRIC = "decoded frames/s (NMS 12 it + order-2 OSD on NMS failures, (128,64) CCSDS, Eb/N0 2.5 dB)"
# algorithmic HBM bytes per frame: NMS reads 512 B LLR, writes 16 B bits + 2 status bytes;
# OSD reads 4 B index + 512 B LLR, writes 16 B codeword (+ 4 B TEP index when requested)
NMS_BYTES = 512 + 16 + 2
OSD_BYTES = 4 + 512 + 16


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--frames", type=int, default=1 << 21, help="frames per GPU per step")
    p.add_argument("--e2e-frames", type=int, default=1 << 21)
    p.add_argument("--ebn0", type=float, default=2.5)
    p.add_argument("--order", type=int, default=2)
    p.add_argument("--cpu-sample", type=int, default=40000, help="frames of the CPU baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    return p.parse_args()


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def cpu_pipeline_sample(code, n_frames, order, ebn0, threads=0, repeat=1):
    """The CPU arm: C port of the reference path (oracle/c/ldpc_oracle.c) on all host threads:
    NMS 12 it on n_frames, order-p OSD on the detected failures.  Returns frames/s and details."""
    from oracle import c_oracle as CO
    from oracle import osd_oracle as OO
    from oracle import philox_oracle as PO

    y, cw, _ = PO.gen_frames(123, 0, n_frames, ebn0, code.G)
    teps = OO.pack_teps(OO.generate_teps_conv(order))
    best = None
    for _ in range(repeat):
        t0 = time.perf_counter()
        r = CO.nms(y, code.H, 12, ALPHA, threads=threads)
        t1 = time.perf_counter()
        fails = np.flatnonzero(r["syndrome_nz"])
        o = CO.osd(np.ascontiguousarray(y[fails]), None, code.G, teps, threads=threads, want_perm=False)
        t2 = time.perf_counter()
        dt = t2 - t0
        if best is None or dt < best[0]:
            best = (dt, t1 - t0, t2 - t1, len(fails))
    dt, t_nms, t_osd, nf = best
    return {"value": n_frames / dt, "nms_frames_per_s": n_frames / t_nms, "osd_frames_per_s": nf / max(t_osd, 1e-9),
            "seconds": dt, "failed": nf, "cores": CO.max_threads() if threads <= 0 else threads}


# ---------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from short_ldpc_decoding_osd_b200.fill_matrix_info import Code

    code = Code()
    from oracle import c_oracle as CO

    CO.build()
    per_step = args.cpu_sample
    for _ in range(max(args.warmup, 1)):
        cpu_pipeline_sample(code, min(per_step, 5000), args.order, args.ebn0)
    t0 = time.perf_counter()
    dets = [cpu_pipeline_sample(code, per_step, args.order, args.ebn0) for _ in range(args.steps)]
    secs = sum(d["seconds"] for d in dets)
    value = per_step * args.steps / secs
    cores = dets[0]["cores"]
    sample = f"{per_step} frames/step x {args.steps} steps of the same workload (NMS on all, OSD-{args.order} on the ~{dets[0]['failed']} detected failures)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32/i64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": per_step, "osd_order": args.order, "ebn0_db": args.ebn0,
                   "note": "TensorFlow is not installable offline; the reference's TF-eager path is timed through its C port (oracle/c), OpenMP over all host threads"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample,
                         "nms_frames_per_s": statistics.median(d["nms_frames_per_s"] for d in dets),
                         "osd_frames_per_s": statistics.median(d["osd_frames_per_s"] for d in dets)},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from short_ldpc_decoding_osd_b200 import _lib
    from short_ldpc_decoding_osd_b200.fill_matrix_info import Code

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    devs = f"cuda:{local_rank}"
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner with printf when the communicator is
        # created, so fd 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(devs))
            warm = torch.zeros(1, device=devs)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    code = Code()
    h = _lib.Handle(code.H, code.G, device=local_rank)
    B, K, W, order = args.frames, args.steps, max(args.warmup, 3), args.order
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    # ---- synthetic batch, resident in HBM: rank r owns frames [r*B, (r+1)*B) of the run ----------
    llr = torch.empty((B, 128), dtype=torch.float32, device=devs)
    truth = torch.empty((B, 4), dtype=torch.int32, device=devs)
    bits = torch.empty((B, 4), dtype=torch.int32, device=devs)
    syn = torch.empty((B,), dtype=torch.uint8, device=devs)
    counters = torch.zeros(16, dtype=torch.int64, device=devs)
    h.call("ldpcb_gen_frames", 2024, rank * B, B, args.ebn0, llr, truth, sp)

    def step():
        counters.zero_()
        h.call("ldpcb_decode", llr, B, 12, ALPHA, 1.0, 1.0, 0, order, 0, bits, syn, None, truth, counters, sp)
        if world > 1:
            dist.all_reduce(counters)  # the only collective of the path: 128 bytes of tallies

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = h.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    ev[0].record(stream)
    for i in range(K):
        step()
        ev[i + 1].record(stream)
    barrier()
    total_ms = ev[0].elapsed_time(ev[K])
    launches = h.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([total_ms], dtype=torch.float64, device=devs)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * B * K / (total_ms * 1e-3)
    cnt = counters.cpu().numpy()
    frames_all = int(cnt[0])

    # ---- per-kernel times for the roofline (same inputs, CUDA events on the launching stream) -----
    def timed(fn, n=5):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.mean(ts)

    it = torch.empty((B,), dtype=torch.uint8, device=devs)
    nms_ms = timed(lambda: h.call("ldpcb_nms_decode", llr, B, 12, ALPHA, 1.0, 1.0, 0, bits, it, syn, None, sp))
    idx = torch.empty((B,), dtype=torch.int32, device=devs)
    nfail_t = torch.empty((1,), dtype=torch.int32, device=devs)
    h.call("ldpcb_select_flagged", syn, B, idx, nfail_t, sp)
    nfail = int(nfail_t.item())
    fl = torch.empty((max(nfail, 1), 128), dtype=torch.float32, device=devs)
    h.call("ldpcb_gather_rows", llr, idx, nfail_t, B, 128, fl, sp)
    cwb = torch.empty((max(nfail, 1), 4), dtype=torch.int32, device=devs)
    osd_ms = timed(lambda: h.call("ldpcb_osd_decode", fl, fl, nfail, order, 0, 0, cwb, None, None, None, None, None, sp)) if nfail else 0.0
    peaks, peak_src = measured_peaks()
    traffic, prof = None, {}
    try:  # per-frame DRAM bytes and warp instructions measured once under ncu (profiles/), scaled to this launch
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            prof = json.load(f)
    except Exception:
        prof = {}
    step_ms = total_ms / K
    dom_is_osd = osd_ms >= nms_ms
    dom_key = "osd_kernel" if dom_is_osd else "nms_kernel"
    dom_ms = osd_ms if dom_is_osd else nms_ms
    dom_frames = nfail if dom_is_osd else B
    dom_bytes = (OSD_BYTES if dom_is_osd else NMS_BYTES) * dom_frames
    if dom_key in prof:
        traffic = prof[dom_key]["dram_bytes_per_frame"] * dom_frames
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    sm_clock_hz = 1e6 * float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
    issue_peak = h.sm_count * 4 * sm_clock_hz  # one warp instruction per scheduler (4 per SM) per cycle

    def issue(key, frames, ms):
        if key not in prof or "warp_instr_per_frame" not in prof[key] or not ms:
            return None
        rate = prof[key]["warp_instr_per_frame"] * frames / (ms * 1e-3)
        return {"warp_instr_per_frame": prof[key]["warp_instr_per_frame"], "achieved_warp_instr_per_s": rate,
                "peak_warp_instr_per_s": issue_peak, "frac": rate / issue_peak}

    roofline = {
        "bound": "hbm", "kernel": "osd_pair_kernel (warp-local tensor-core pair sweep)" if dom_is_osd else "nms_kernel<5,3,true,false,false>",
        "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
        "traffic": traffic, "traffic_source": "profiles/r01_traffic.json (ncu dram__bytes per frame x frames of this launch)" if traffic else None,
        "peak_source": peak_src, "kernel_ms": dom_ms, "share_of_step": dom_ms / step_ms,
        "algorithmic_bytes_per_launch": dom_bytes,
        "note": "both kernels are integer/FP32 issue-bound, not HBM-bound (SURVEY.md 8d): the HBM fraction is reported as BASELINE asks; "
                "the binding resource is SM issue slots -- `issue` gives warp instructions per frame (ncu sm__inst_executed, profiles/) x "
                "frames/s of this run against SMs x 4 schedulers x the SM clock sampled in this run",
        "issue": {"nms": issue("nms_kernel", B, nms_ms), "osd": issue("osd_kernel", nfail, osd_ms)},
        "kernels": {"nms_ms": nms_ms, "nms_frames_per_s": B / (nms_ms * 1e-3), "nms_GBps": NMS_BYTES * B / (nms_ms * 1e-3) / 1e9,
                    "osd_ms": osd_ms, "osd_frames": nfail, "osd_frames_per_s": (nfail / (osd_ms * 1e-3)) if osd_ms else None,
                    "osd_GBps": (OSD_BYTES * nfail / (osd_ms * 1e-3) / 1e9) if osd_ms else None},
    }

    # ---- end to end through the host-buffer C-ABI call -------------------------------------------
    Be = min(args.e2e_frames, B)
    yh = _lib.pinned_empty((Be, 128), np.float32)
    th = _lib.pinned_empty((Be, 4), np.uint32)
    bh = _lib.pinned_empty((Be, 4), np.uint32)
    sh = _lib.pinned_empty((Be,), np.uint8)
    yh[:] = llr[:Be].cpu().numpy()
    th[:] = truth[:Be].cpu().numpy().view(np.uint32)
    ch = np.zeros(16, np.uint64)

    def e2e_step():
        ch[:] = 0
        h.call("ldpcb_decode_host", yh, Be, 12, ALPHA, 1.0, 1.0, 0, order, 0, bh, sh, None, th, ch)

    for _ in range(2):
        e2e_step()
    barrier()
    e2e_steps = max(3, min(K, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=devs)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * Be * e2e_steps / float(t.item())
    e2e = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": Be * (512 + 16), "d2h_bytes_per_step": Be * (16 + 1) + 128,
           "frames_per_step": Be, "steps": e2e_steps, "api": "ldpcb_decode_host (pinned host LLRs + truth bits in, decisions + syndrome flags + counters out)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle as CO

        CO.build()
        d = cpu_pipeline_sample(code, args.cpu_sample, order, args.ebn0, repeat=2)
        cpu = {"value": d["value"], "unit": "frames/s", "cores": d["cores"], "kind": "port",
               "sample": f"{args.cpu_sample} frames of the same workload (C port of the reference path, OpenMP; NMS on all, OSD-{order} on {d['failed']} failures), best of 2",
               "nms_frames_per_s": d["nms_frames_per_s"], "osd_frames_per_s": d["osd_frames_per_s"]}

    if rank == 0:
        fer_nms = cnt[1] / max(frames_all, 1)
        fer_final = cnt[9] / max(frames_all, 1)
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/i64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": B, "osd_order": order, "tep_count": h.tep_count(order),
                       "ebn0_db": args.ebn0, "nms_iters": 12, "early_stop": 0, "alpha": ALPHA,
                       "l2": "inputs (1 GiB LLR per step) larger than L2, no flush", "parallelism": f"frames sharded over {world} GPU(s), counter all-reduce only"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "tallies": {"frames": frames_all, "fer_nms": fer_nms, "fer_after_osd": fer_final, "osd_frames": int(cnt[6]),
                        "undetected_nms": int(cnt[4])},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    h.close()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
