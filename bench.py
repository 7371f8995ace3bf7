#!/usr/bin/env python
"""Headline benchmark: decoded frames/s of the hot path (NMS 12 iterations + order-2 OSD on the NMS
failures, CCSDS (128,64), Eb/N0 = 2.5 dB) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W             # our arm, one JSON line on rank 0
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the reference's own path

A step = one pass of the whole device pipeline (ldpcb_decode: NMS on all frames, compaction of the
detected failures, order-2 OSD on them, tallies) over one batch of B synthetic frames per GPU that is
already resident in HBM (generated once by the Philox kernel), followed by the all-reduce of the 16
uint64 counters.  The batch (B * 512 B = 1 GiB at the default B = 2^21) is larger than L2 (126 MB), so
no explicit flush is needed between steps.  `e2e` is the same pipeline through the host-buffer C-ABI
call (ldpcb_decode_host: pinned host LLRs in, decisions + counters out, copies inside the timed region).

The JSON line also carries
  roofline   the binding bound of the dominant kernel (SM issue slots, algorithmic warp instructions per
             frame derived in DESIGN.md 6.1) with the executed-instruction utilisation and the HBM
             fraction beside it
  configs    CUDA-event rates of every kernel family on the same resident batch (NMS with and without
             early stop, OSD order 0..3 on the NMS failures, FS, PB at 3.0 dB, the DL scheme, the fused
             Monte-Carlo step, the frame generator)
  cpu_baseline  the reference path on the host cores (N = 1 only)

The reference arm times, in this order of preference: the unmodified reference under a real TensorFlow
(kind "tf": needs `import tensorflow` and the reference tree at baseline/_ref or $LDPCB_REFERENCE_ROOT),
else the C/OpenMP port of the same path (kind "port", oracle/c) on every host core -- explicitly, since
torchrun exports OMP_NUM_THREADS=1.
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
# algorithmic warp instructions per frame (32-lane operations the arithmetic of the path needs,
# independent of how a kernel is written)
NMS_ALG_WARP_INSTR = 2038
OSD2_ALG_WARP_INSTR = 1800
# SURVEY.md 8d's first estimates of the same quantities (12 lane-ops per edge update; 1.7e5 lane-ops per order-2 frame),
# kept beside the derived counts so that round-1 fractions stay comparable
NMS_SURVEY_WARP_INSTR = 2438
OSD2_SURVEY_WARP_INSTR = 5312
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_traffic.json")


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
    p.add_argument("--no-configs", action="store_true", help="skip the per-kernel `configs` section")
    return p.parse_args()


def host_cores() -> int:
    """Cores this process may run on (torchrun's OMP_NUM_THREADS=1 is deliberately ignored)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


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


def load_traffic():
    """Per-frame DRAM bytes and executed warp instructions captured under ncu by scripts/make_traffic.py.  The file
    is stamped with the hash of the kernel sources it was measured on; a file from another build is refused."""
    from short_ldpc_decoding_osd_b200 import build as B

    try:
        with open(TRAFFIC_FILE) as f:
            prof = json.load(f)
    except Exception:
        return {}, "no profiles/r02_traffic.json"
    if prof.get("build_stamp") != B._stamp():
        return {}, "profiles/r02_traffic.json was captured on another build of the kernels (stamp mismatch): refused"
    return prof, "profiles/r02_traffic.json (ncu, stamp matches this build)"


def cpu_pipeline_sample(code, n_frames, order, ebn0, threads=0, repeat=1, seed=2024):
    """The CPU arm: C port of the reference path (oracle/c/ldpc_oracle.c) on `threads` host threads:
    NMS 12 it on n_frames, order-p OSD on the detected failures.  The frames are the first n_frames of the
    run the GPU arm decodes (same Philox seed and counters).  Returns frames/s and details."""
    from oracle import c_oracle as CO
    from oracle import osd_oracle as OO
    from oracle import philox_oracle as PO

    threads = threads if threads > 0 else host_cores()
    y, cw, _ = PO.gen_frames(seed, 0, n_frames, ebn0, code.G)
    teps = OO.pack_teps(OO.generate_teps_conv(order))
    best = None
    for _ in range(repeat):
        t0 = time.perf_counter()
        r = CO.nms(y, code.H, 12, ALPHA, threads=threads)
        t1 = time.perf_counter()
        fails = np.flatnonzero(r["syndrome_nz"])
        CO.osd(np.ascontiguousarray(y[fails]), None, code.G, teps, threads=threads, want_perm=False)
        t2 = time.perf_counter()
        dt = t2 - t0
        if best is None or dt < best[0]:
            best = (dt, t1 - t0, t2 - t1, len(fails))
    dt, t_nms, t_osd, nf = best
    return {"value": n_frames / dt, "nms_frames_per_s": n_frames / t_nms, "osd_frames_per_s": nf / max(t_osd, 1e-9),
            "seconds": dt, "failed": nf, "cores": threads}


# ---------------------------------------------------------------------------------------------------
def reference_root():
    """Where an unmodified copy of the reference lives at run time (never /root/reference: it does not exist on
    the GPU box)."""
    for cand in (os.environ.get("LDPCB_REFERENCE_ROOT"), os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "LDPC_128", "Ldpc_128_testing")):
            return cand
    return None


def try_tf_reference(args):
    """if TensorFlow imports and the reference tree is present, time the UNMODIFIED reference
    (Decoding_model in batches of 1000, swapped_info + convention_osd_main per failure) in a child process.
    -> (dict, None) or (None, why-not)."""
    root = reference_root()
    if root is None:
        return None, "no reference tree at baseline/_ref or $LDPCB_REFERENCE_ROOT"
    probe = subprocess.run([sys.executable, "-c", "import tensorflow as tf; print(tf.__file__)"], capture_output=True, text=True)
    if probe.returncode != 0:
        return None, "import tensorflow failed: " + (probe.stderr.strip().splitlines() or ["?"])[-1][:160]
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)  # let TF use every core
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "tf_reference_arm.py"), "--root", root, "--frames", str(args.tf_frames),
           "--order", str(args.order), "--ebn0", str(args.ebn0), "--steps", str(args.steps), "--warmup", str(args.warmup)]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=1500)
    if r.returncode != 0:
        return None, "reference under TensorFlow failed: " + (r.stderr.strip().splitlines() or ["?"])[-1][:200]
    return json.loads(r.stdout.strip().splitlines()[-1]), None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun rank 0 alone runs the CPU arm
    from short_ldpc_decoding_osd_b200.fill_matrix_info import Code

    args.tf_frames = int(os.environ.get("LDPCB_TF_FRAMES", "500"))
    cores = host_cores()
    tf_res, why_not_tf = try_tf_reference(args)
    base = {
        "impl": "reference", "metric": METRIC, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/i64", "data": "synthetic", "gpu_launches": 0,
    }
    if tf_res is not None and tf_res.get("kind", "tf") != "tf" and os.environ.get("LDPCB_ALLOW_TF_SHIM") != "1":
        tf_res, why_not_tf = None, "`import tensorflow` resolved to the NumPy shim under oracle/tf_shim, which is test infrastructure, not TensorFlow"
    if tf_res is not None:
        value = tf_res["value"]
        line = dict(base, value=value, ms_per_step=tf_res["ms_per_step"],
                    config={"workload": WORKLOAD, "frames_per_step": tf_res["frames_per_step"], "osd_order": args.order, "ebn0_db": args.ebn0,
                            "note": "unmodified reference under TensorFlow: ms_test.Decoding_model (B=1000) + swapped_info + convention_osd_main per NMS failure"},
                    cpu_baseline={"value": value, "unit": "frames/s", "cores": cores, "kind": tf_res.get("kind", "tf"), "sample": tf_res["sample"],
                                  "nms_frames_per_s": tf_res.get("nms_frames_per_s"), "osd_frames_per_s": tf_res.get("osd_frames_per_s")},
                    e2e={"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line), flush=True)
        return
    code = Code()
    from oracle import c_oracle as CO

    CO.build()
    per_step = args.cpu_sample
    for _ in range(max(args.warmup, 1)):
        cpu_pipeline_sample(code, min(per_step, 5000), args.order, args.ebn0, threads=cores)
    dets = [cpu_pipeline_sample(code, per_step, args.order, args.ebn0, threads=cores) for _ in range(args.steps)]
    secs = sum(d["seconds"] for d in dets)
    value = per_step * args.steps / secs
    sample = (f"{per_step} frames/step x {args.steps} steps: the first {per_step} frames of the GPU arm's run (same Philox seed), "
              f"NMS on all, OSD-{args.order} on the ~{dets[0]['failed']} detected failures")
    line = dict(base, value=value, ms_per_step=1e3 * secs / args.steps,
                config={"workload": WORKLOAD, "frames_per_step": per_step, "osd_order": args.order, "ebn0_db": args.ebn0,
                        "note": "the reference's TF-eager path timed through its C port (oracle/c), OpenMP over all host cores; TF path not taken: " + str(why_not_tf)},
                cpu_baseline={"value": value, "unit": "frames/s", "cores": dets[0]["cores"], "kind": "port", "sample": sample,
                              "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
                              "nms_frames_per_s": statistics.median(d["nms_frames_per_s"] for d in dets),
                              "osd_frames_per_s": statistics.median(d["osd_frames_per_s"] for d in dets)},
                e2e={"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
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
    # host side of the e2e path: this rank's threads and pinned buffers go to the NUMA node of its GPU
    all_cpus = os.sched_getaffinity(0)
    numa = _lib.bind_host_to_device(local_rank)
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
    def timed(fn, n=5, warm=1):
        for _ in range(warm):
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
    prof, prof_src = load_traffic()
    step_ms = total_ms / K
    sm_clock_hz = 1e6 * float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
    issue_peak = h.sm_count * 4 * sm_clock_hz  # one warp instruction per scheduler (4 per SM) per cycle

    def kernel_roof(key, alg_instr, survey_instr, alg_bytes, frames, ms):
        if not ms or not frames:
            return None
        rate = frames / (ms * 1e-3)
        d = {"ms": ms, "frames": frames, "frames_per_s": rate, "share_of_step": ms / step_ms,
             "issue": {"algorithmic_warp_instr_per_frame": alg_instr, "achieved_warp_instr_per_s": alg_instr * rate,
                       "peak_warp_instr_per_s": issue_peak, "frac": alg_instr * rate / issue_peak,
                       "survey_8d_warp_instr_per_frame": survey_instr, "frac_with_survey_count": survey_instr * rate / issue_peak},
             "hbm": {"algorithmic_bytes_per_frame": alg_bytes, "achieved_GBps": alg_bytes * rate / 1e9, "peak_GBps": peaks["hbm_gbs"],
                     "frac": alg_bytes * rate / 1e9 / peaks["hbm_gbs"]}}
        p = prof.get(key)
        if p:
            d["executed"] = {"warp_instr_per_frame": p["warp_instr_per_frame"], "issue_slot_utilisation": p["warp_instr_per_frame"] * rate / issue_peak,
                             "dram_bytes_per_frame": p["dram_bytes_per_frame"]}
        return d

    k_nms = kernel_roof("nms_kernel", NMS_ALG_WARP_INSTR, NMS_SURVEY_WARP_INSTR, NMS_BYTES, B, nms_ms)
    k_osd = kernel_roof("osd_kernel", OSD2_ALG_WARP_INSTR, OSD2_SURVEY_WARP_INSTR, OSD_BYTES, nfail, osd_ms)
    dom, dom_name = (k_osd, "osd_pair_kernel (order 2, warp-local tensor-core pair sweep)") if osd_ms >= nms_ms else (k_nms, "nms kernel (12 iterations, fixed)")
    dom_exec = (dom or {}).get("executed")
    roofline = {
        "bound": "issue", "kernel": dom_name,
        "achieved": dom["issue"]["achieved_warp_instr_per_s"], "peak": issue_peak, "unit": "warp-instr/s", "frac": dom["issue"]["frac"],
        "traffic": (dom_exec["dram_bytes_per_frame"] * dom["frames"]) if dom_exec else None,
        "traffic_source": prof_src, "peak_source": f"{h.sm_count} SMs x 4 schedulers x {sm_clock_hz / 1e6:.0f} MHz sampled in this run",
        "kernel_ms": dom["ms"], "share_of_step": dom["share_of_step"],
        "algorithmic_warp_instr_per_launch": dom["issue"]["algorithmic_warp_instr_per_frame"] * dom["frames"],
        "algorithmic_bytes_per_launch": dom["hbm"]["algorithmic_bytes_per_frame"] * dom["frames"],
        "hbm": {"achieved": dom["hbm"]["achieved_GBps"], "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": dom["hbm"]["frac"], "peak_source": peak_src},
        "executed": dom_exec,
        "note": "both decoders are integer/FP32 issue-bound, not HBM-bound (SURVEY.md 8d, traffic == algorithmic bytes): `frac` is ALGORITHMIC warp "
                "instructions per frame (DESIGN.md 6.1) x frames/s over SMs x 4 x the SM clock sampled in this run; `executed` is what the shipped "
                "binary issues (ncu, efficiency = algorithmic / executed); `hbm` is the non-binding bound BASELINE asks for",
        "kernels": {"nms": k_nms, "osd": k_osd},
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
    h2d_b, d2h_b = Be * (512 + 16), Be * (16 + 1) + 128

    # plain pinned H2D copies of the same sizes, all ranks at once: the ceiling the e2e number can reach on this host
    def h2d_ceiling():
        dst, dst2 = torch.empty((Be, 128), dtype=torch.float32, device=devs), torch.empty((Be, 4), dtype=torch.int32, device=devs)
        src, src2 = torch.from_numpy(yh), torch.from_numpy(th.view(np.int32))
        for _ in range(2):
            dst.copy_(src, non_blocking=True); dst2.copy_(src2, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            dst.copy_(src, non_blocking=True); dst2.copy_(src2, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=devs)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return world * Be * (512 + 16) * e2e_steps / float(tt.item()) / 1e9

    ceil_gbs = h2d_ceiling()
    ceil_frames = ceil_gbs * 1e9 / (512.0 + 16.0)
    e2e = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
           "frames_per_step": Be, "steps": e2e_steps, "h2d_GBps": e2e_value * (512 + 16) / 1e9,
           "h2d_ceiling_GBps": ceil_gbs, "h2d_ceiling_frames_per_s": ceil_frames, "frac_of_h2d_ceiling": e2e_value / ceil_frames,
           "ceiling_how": f"plain pinned cudaMemcpyAsync H2D of the same {Be * 528} bytes per step (LLRs + truth words) from all {world} rank(s) at once, same steps, nothing else running",
           "numa": numa,
           "api": "ldpcb_decode_host (pinned host LLRs + truth bits in, decisions + syndrome flags + counters out)"}

    configs = None
    if not args.no_configs:
        try:
            configs = measure_configs(h, torch, llr, truth, fl, nfail, sp, stream, args.ebn0, rank * B)
        except Exception as e:  # the headline must survive a failure of a side measurement
            configs = {"error": repr(e)[:300]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle as CO

        CO.build()
        os.sched_setaffinity(0, all_cpus)  # the CPU arm gets every core again
        d = cpu_pipeline_sample(code, args.cpu_sample, order, args.ebn0, threads=len(all_cpus), repeat=2)
        cpu = {"value": d["value"], "unit": "frames/s", "cores": d["cores"], "kind": "port",
               "sample": f"the first {args.cpu_sample} frames of this run's batch (same Philox seed 2024; C port of the reference path, OpenMP; NMS on all, OSD-{order} on {d['failed']} failures), best of 2",
               "nms_frames_per_s": d["nms_frames_per_s"], "osd_frames_per_s": d["osd_frames_per_s"]}

    if rank == 0:
        fer_nms = cnt[1] / max(frames_all, 1)
        fer_final = cnt[9] / max(frames_all, 1)
        l2_note = (f"inputs ({B * 512 / 2**20:.0f} MiB LLR per step) larger than L2 (126 MB), no flush" if B * 512 > 2 * 126e6
                   else f"inputs ({B * 512 / 2**20:.0f} MiB LLR per step) NOT larger than L2: L2-resident rates, not a headline configuration")
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/i64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": B, "osd_order": order, "tep_count": h.tep_count(order),
                       "ebn0_db": args.ebn0, "nms_iters": 12, "early_stop": 0, "alpha": ALPHA,
                       "l2": l2_note, "parallelism": f"frames sharded over {world} GPU(s), counter all-reduce only"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "configs": configs,
            "tallies": {"frames": frames_all, "fer_nms": fer_nms, "fer_after_osd": fer_final, "osd_frames": int(cnt[6]),
                        "undetected_nms": int(cnt[4])},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    h.close()


def measure_configs(h, torch, llr, truth, fails_llr, nfail, sp, stream, ebn0, first_frame):
    """CUDA-event rates (frames/s = frames handed to the kernel / mean launch time, 3 timed launches after 1 warm-up)
    of every kernel family of SURVEY 8 on this rank's resident batch.  Sizes are bounded so the section costs seconds."""
    from short_ldpc_decoding_osd_b200 import _lib

    dev = llr.device
    B = llr.shape[0]
    e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)  # noqa: E731

    def rate(frames, fn, n=3):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = statistics.mean(ts)
        return {"frames": int(frames), "ms": ms, "frames_per_s": frames / (ms * 1e-3)}

    out = {}
    bits, it, syn = e((B, 4), torch.int32), e((B,), torch.uint8), e((B,), torch.uint8)
    cnt = torch.zeros(16, dtype=torch.int64, device=dev)
    for early in (0, 1):
        out[f"nms12_early_stop_{early}"] = rate(B, lambda: h.call("ldpcb_nms_decode", llr, B, 12, ALPHA, 1.0, 1.0, early, bits, it, syn, None, sp))
        if early:
            out["nms12_early_stop_1"]["mean_iterations"] = float(it.float().mean().item())
    out["nms12_only_pipeline"] = rate(B, lambda: h.call("ldpcb_decode", llr, B, 12, ALPHA, 1.0, 1.0, 0, -1, 0, bits, syn, None, truth, cnt, sp))
    out["nms12_early_stop_1_plus_osd2_pipeline"] = rate(B, lambda: h.call("ldpcb_decode", llr, B, 12, ALPHA, 1.0, 1.0, 1, 2, 0, bits, syn, None, truth, cnt, sp))
    cw = e((max(nfail, 1), 4), torch.int32)
    for order in (0, 1, 2, 3):
        n = nfail
        out[f"osd_order{order}_on_nms_failures"] = rate(n, lambda: h.call("ldpcb_osd_decode", fails_llr, fails_llr, n, order, 0, 0, cw, None, None, None, None, None, sp))
    nt, sk = e((max(nfail, 1),), torch.int32), e((max(nfail, 1),), torch.uint8)
    for order in (2, 3):
        n = min(nfail, 1 << 17)
        out[f"fs_osd_order{order}_on_nms_failures"] = rate(n, lambda: h.call("ldpcb_osd_fs_decode", fails_llr, n, order, 6.5, 30, 6.4, cw, None, nt, sk, None, None, None, sp))
        out[f"fs_osd_order{order}_on_nms_failures"]["mean_teps"] = float(nt[:n].float().mean().item())
    # PB-OSD: BASELINE config 3 is order 2 at 3.0 dB; order 3 is the reference's default order_limit
    Bp = min(B, 1 << 20)
    y3, t3, b3, s3, i3 = e((Bp, 128), torch.float32), e((Bp, 4), torch.int32), e((Bp, 4), torch.int32), e((Bp,), torch.uint8), e((Bp,), torch.uint8)
    h.call("ldpcb_gen_frames", 2025, first_frame, Bp, 3.0, y3, t3, sp)
    h.call("ldpcb_nms_decode", y3, Bp, 12, ALPHA, 1.0, 1.0, 0, b3, i3, s3, None, sp)
    f3 = y3[s3.bool()].contiguous()
    st4 = e((max(f3.shape[0], 1), 4), torch.int32)
    for order, cap in ((2, 1 << 17), (3, 1 << 16)):
        n = min(f3.shape[0], cap)
        if n:
            out[f"pb_osd_order{order}_ebn0_3.0dB_on_nms_failures"] = rate(n, lambda: h.call("ldpcb_osd_pb_decode", f3, n, order, 3.0, cw, st4, None, None, sp))
            out[f"pb_osd_order{order}_ebn0_3.0dB_on_nms_failures"]["mean_teps"] = float(st4[:n, 0].float().mean().item())
    out["nms12_plus_osd2_pipeline_ebn0_3.0dB"] = rate(Bp, lambda: h.call("ldpcb_decode", y3, Bp, 12, ALPHA, 1.0, 1.0, 0, 2, 0, b3, s3, None, t3, cnt, sp))
    # Monte-Carlo step (generator + decode, nothing resident) and the generator alone
    Bs = min(B, 1 << 20)
    out["gen_frames"] = rate(Bs, lambda: h.call("ldpcb_gen_frames", 7, first_frame, Bs, ebn0, llr, truth, sp))
    h.call("ldpcb_gen_frames", 2024, first_frame, Bs, ebn0, llr, truth, sp)  # restore the batch
    out["simulate_gen_plus_nms12_plus_osd2"] = rate(Bs, lambda: h.call("ldpcb_simulate", 7, first_frame, Bs, ebn0, 12, ALPHA, 1.0, 1.0, 0, 2, 0, cnt, sp))
    # DL scheme (BASELINE config 4 shape; synthetic taps and window classifier: the trained checkpoints are not shipped)
    try:
        from short_ldpc_decoding_osd_b200 import globalmap as GL, nn_net, nn_testing, simulate
        from short_ldpc_decoding_osd_b200 import ordered_statistics_decoding as OSD
        from short_ldpc_decoding_osd_b200.fill_matrix_info import Code

        code = Code()
        for k, v in dict(code_parameters=code, selected_decoder_type="NMS-1", num_iterations=12, threshold_sum=2, segment_num=6, soft_margin=0.9,
                         decoding_length=30, sliding_win_width=5).items():
            GL.set_map(k, v)
        tep_info = nn_testing.generate_teps(OSD.osd(code), nn_testing.filter_order_patterns(nn_testing.convention_segment_path()))
        rng = np.random.default_rng(3)
        taps = (np.full(13, 1 / 13) + 0.03 * rng.normal(size=13)).astype(np.float32)
        net = nn_net.Predict_outlier_light(5, W1=np.eye(6, dtype=np.float32), W2=np.array([[0, -0.5], [0, 0.5], [0, 0], [0, 0], [0, 0], [0, 0.15]], np.float32))
        n = 1 << 22
        simulate.run_point_dl(h, ebn0, 1 << 22, tep_info, taps, 0.05, net.W1, net.W2, seed=1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        simulate.run_point_dl(h, ebn0, n, tep_info, taps, 0.05, net.W1, net.W2, seed=2)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["dl_scheme_gen_nms_fir_blockminima_window"] = {"frames": n, "ms": 1e3 * dt, "frames_per_s": n / dt, "timing": "wall clock around simulate.run_point_dl (host loop + device)",
                                                         "teps_on_path": int(tep_info[1][-1])}
    except Exception as ex:
        out["dl_scheme_gen_nms_fir_blockminima_window"] = {"error": repr(ex)[:200]}
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
