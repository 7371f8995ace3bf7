"""FER-curve harness: the reference's file-coupled chain Testing_data_gen_128 -> Ldpc_128_testing -> *_OSD
("Training and Testing recipe.txt":9-18) as one in-memory Monte-Carlo loop on 1..8 GPUs.

Frames are independent, so a run is sharded by frame index: rank r of W takes the contiguous range
``shard_range(total, r, W)`` of every chunk, generates those frames with the counter-based Philox kernel,
decodes them (NMS, then OSD on the detected failures) and tallies on the device -- all inside
``ldpcb_simulate``; the only traffic between GPUs is a sum all-reduce of the 16 uint64 counters per chunk
(128 bytes), which also drives the reference's stopping rules (stop a point after N frame errors:
``ldpc_128_testing.py:36,130``; after N OSD failures: ``pb_testing.py:174``).

    torchrun --nproc-per-node 8 -m short_ldpc_decoding_osd_b200.simulate --ebn0 1.5 2.0 2.5 3.0 3.5 4.0 --order 1
"""
from __future__ import annotations

import argparse
import json
import math
import os
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib

CN = {n: i for i, n in enumerate(_lib.COUNTER_NAMES)}


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous frame range [a, b) of `rank`: sizes differ by at most one, ranges tile [0, total)."""
    base, rem = divmod(int(total), int(world))
    a = rank * base + min(rank, rem)
    return a, a + base + (1 if rank < rem else 0)


def allreduce_counters(local: np.ndarray) -> np.ndarray:
    """Sum of the uint64 counter blocks over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist

    local = np.ascontiguousarray(local, dtype=np.uint64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local.copy()
    t = torch.from_numpy(local.view(np.int64).copy())
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().view(np.uint64)


def should_stop(total: np.ndarray, max_frame_errors: Optional[int], counter: str = "final_frame_err") -> bool:
    return max_frame_errors is not None and int(total[CN[counter]]) >= int(max_frame_errors)


def wilson_interval(k: int, n: int, z: float = 1.959963984540054) -> Tuple[float, float]:
    if n == 0:
        return 0.0, 1.0
    p = k / n
    d = 1 + z * z / n
    c = p + z * z / (2 * n)
    h = z * math.sqrt(p * (1 - p) / n + z * z / (4 * n * n))
    return max(0.0, (c - h) / d), min(1.0, (c + h) / d)


def clopper_pearson(k: int, n: int, conf: float = 0.95) -> Tuple[float, float]:
    from scipy.stats import beta

    a = (1 - conf) / 2
    lo = 0.0 if k == 0 else float(beta.ppf(a, k, n - k + 1))
    hi = 1.0 if k == n else float(beta.ppf(1 - a, k + 1, n - k))
    return lo, hi


class Tallies:
    """Named view of a counter block (include/ldpc_b200.h LDPCB_CNT_*)."""

    def __init__(self, counters: np.ndarray, n_bits: int = 128):
        self.c = np.asarray(counters, dtype=np.uint64).copy()
        self.n_bits = n_bits

    def __getattr__(self, name):
        if name in CN:
            return int(self.c[CN[name]])
        raise AttributeError(name)

    @property
    def fer_nms(self):
        return self.nms_frame_err / max(self.frames, 1)

    @property
    def ber_nms(self):
        return self.nms_bit_err / max(self.frames * self.n_bits, 1)

    @property
    def fer_final(self):
        return self.final_frame_err / max(self.frames, 1)

    @property
    def fer_osd(self):
        """FER of the OSD stage alone, on the frames it was given; final FER = FER_NMS(detected) * FER_OSD + undetected."""
        return self.osd_frame_err / max(self.osd_frames, 1)

    def reference_log_line(self) -> str:
        """The line ldpc_128_testing.py:137-140 prints/writes for an SNR point."""
        return "FER %.4f, BER %.4f,UFER %.6f" % (self.fer_nms, self.ber_nms, self.nms_undetected / max(self.frames, 1))

    def as_dict(self):
        d = {n: int(self.c[i]) for n, i in CN.items()}
        d.update(fer_nms=self.fer_nms, fer_final=self.fer_final, fer_osd=self.fer_osd, ber_nms=self.ber_nms,
                 fer_final_ci95=wilson_interval(self.final_frame_err, self.frames), fer_nms_ci95=wilson_interval(self.nms_frame_err, self.frames))
        return d


def run_point(handle, ebn0_db: float, total_frames: int, seed: int = 0, osd_order: int = 2, tep_order: int = _lib.TEP_CONV,
              iters: int = 12, alpha: float = 0.66943514, w_vc: float = 1.0, w_marg: float = 1.0, early_stop: int = 0,
              chunk: int = 1 << 22, max_frame_errors: Optional[int] = None, stop_counter: str = "final_frame_err",
              rank: int = 0, world: int = 1) -> Tallies:
    """One Eb/N0 point.  `chunk` frames (whole job) per all-reduce; every rank runs its shard of each chunk."""
    import torch

    dev = f"cuda:{handle.device}"
    counters = torch.zeros(_lib.NUM_COUNTERS, dtype=torch.int64, device=dev)
    total = np.zeros(_lib.NUM_COUNTERS, dtype=np.uint64)
    stream = torch.cuda.current_stream(dev).cuda_stream
    done = 0
    while done < total_frames:
        n = min(chunk, total_frames - done)
        a, b = shard_range(n, rank, world)
        counters.zero_()
        if b > a:
            handle.call("ldpcb_simulate", int(seed), int(done + a), int(b - a), float(ebn0_db), int(iters), float(alpha), float(w_vc),
                        float(w_marg), int(early_stop), int(osd_order), int(tep_order), counters, stream)
        total += allreduce_counters(counters.cpu().numpy().view(np.uint64))
        done += n
        if should_stop(total, max_frame_errors, stop_counter):
            break
    return Tallies(total)


def run_point_dl(handle, ebn0_db: float, total_frames: int, tep_info, taps, bias: float, W1, W2, soft_margin: float = 0.9,
                 win_width: int = 5, seed: int = 0, iters: int = 12, alpha: float = 0.66943514, chunk: int = 1 << 22,
                 rank: int = 0, world: int = 1):
    """One Eb/N0 point of the DL scheme (BASELINE config 4), everything on the device: generate -> NMS with the DIA FIR
    (ordering metric) fused in -> detected failures -> block minima along the decoding path, scored against the
    channel LLR -> sliding-window policy.  Mirrors Ldpc_128_testing followed by
    DL_OSD_Testing_serial/nn_testing.Testing_OSD (:159-256).  tep_info = (list of int[T_b,64] blocks over DL MRB
    indices, cumulative sizes) as nn_testing.generate_teps returns.  -> (Tallies of the NMS stage, dict of DL sums)."""
    import torch

    from .ordered_statistics_decoding import FLAGS_DL, pack_dl_teps

    dev = f"cuda:{handle.device}"
    stream = torch.cuda.current_stream(dev).cuda_stream
    teps_list, acc = tep_info
    packed = torch.from_numpy(np.concatenate([pack_dl_teps(b) for b in teps_list]).view(np.int32)).to(dev)
    flags_dl = FLAGS_DL | (min(4, max(1, max(int(np.asarray(b).sum(axis=1).max()) for b in teps_list))) << _lib.OSD_MAXW_SHIFT)
    starts = torch.from_numpy(np.asarray(acc, dtype=np.int32)).to(dev)
    nb = len(teps_list)
    taps = np.ascontiguousarray(taps, dtype=np.float32)
    rows = iters + 1
    W1 = np.ascontiguousarray(W1, dtype=np.float32)
    W2 = np.ascontiguousarray(W2, dtype=np.float32)
    acc_np = np.ascontiguousarray(acc, dtype=np.int32)
    counters = torch.zeros(_lib.NUM_COUNTERS, dtype=torch.int64, device=dev)
    dl = torch.zeros(4, dtype=torch.int64, device=dev)
    tot = np.zeros(_lib.NUM_COUNTERS, dtype=np.uint64)
    tot_dl = np.zeros(4, dtype=np.uint64)
    done = 0
    e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)  # noqa: E731
    while done < total_frames:
        n = min(chunk, total_frames - done)
        a, b = shard_range(n, rank, world)
        m = b - a
        counters.zero_()
        dl.zero_()
        if m > 0:
            llr, truth, bits = e((m, 128), torch.float32), e((m, 4), torch.int32), e((m, 4), torch.int32)
            its, syn = e((m,), torch.uint8), e((m,), torch.uint8)
            idx, cnt = e((m,), torch.int32), e((1,), torch.int32)
            handle.call("ldpcb_gen_frames", int(seed), int(done + a), m, float(ebn0_db), llr, truth, stream)
            # one NMS pass with the DIA FIR fused in: the ordering metric of every frame costs four multiply-adds per
            # iteration and 512 B of HBM, far less than re-decoding the detected failures for their trajectories
            metric_all = e((m, 128), torch.float32)
            its.fill_(iters)
            handle.call("ldpcb_nms_decode_fir", llr, m, iters, float(alpha), 1.0, 1.0, taps, float(bias), bits, syn, metric_all, stream)
            handle.call("ldpcb_tally", bits, syn, its, None, None, -1, 0, truth, m, counters, stream)
            handle.call("ldpcb_select_flagged", syn, m, idx, cnt, stream)
            nf = int(cnt.item())
            if nf > 0:
                llr_f, truth_f, metric = e((nf, 128), torch.float32), e((nf, 4), torch.int32), e((nf, 128), torch.float32)
                handle.call("ldpcb_gather_rows", llr, idx, cnt, nf, 128, llr_f, stream)
                handle.call("ldpcb_gather_rows", metric_all, idx, cnt, nf, 128, metric, stream)
                handle.call("ldpcb_gather_rows", truth.view(torch.float32), idx, cnt, nf, 4, truth_f.view(torch.float32), stream)
                bm, ex, ts = e((nf, nb), torch.int64), e((nf,), torch.int32), e((nf,), torch.int64)
                handle.call("ldpcb_osd_block_minima", metric, llr_f, nf, packed, int(packed.numel()), starts, nb, flags_dl, bm, None, ex,
                            truth_f, ts, None, stream)
                handle.call("ldpcb_dl_window_policy", bm, ex, ts, nf, nb, int(win_width), W1, W2, float(soft_margin), acc_np, None, None, None,
                            dl, stream)
        both = np.concatenate([counters.cpu().numpy().view(np.uint64), dl.cpu().numpy().view(np.uint64), np.zeros(12, np.uint64)])
        red = allreduce_counters(both[:16]), allreduce_counters(both[16:32])
        tot += red[0]
        tot_dl += red[1][:4]
        done += n
    t = Tallies(tot)
    frames = max(t.frames, 1)
    out = {"dl_success": int(tot_dl[0]), "dl_failure": int(tot_dl[1]), "windows_sum": int(tot_dl[2]), "complexity_sum": int(tot_dl[3]),
           "fer_dl_stage": int(tot_dl[1]) / max(int(tot_dl[0] + tot_dl[1]), 1),
           "fer_final": (int(tot_dl[1]) + t.nms_undetected) / frames,
           "avg_teps": int(tot_dl[3]) / max(int(tot_dl[0] + tot_dl[1]), 1)}
    return t, out


def main(argv: Optional[Sequence[str]] = None) -> None:
    import torch
    import torch.distributed as dist

    from .fill_matrix_info import Code

    p = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    p.add_argument("--ebn0", type=float, nargs="+", default=[1.5, 2.0, 2.5, 3.0, 3.5, 4.0])
    p.add_argument("--frames", type=int, default=10_000_000)
    p.add_argument("--order", type=int, default=1)
    p.add_argument("--max-errors", type=int, default=None, help="stop a point after this many residual frame errors")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--early-stop", type=int, default=0)
    p.add_argument("--chunk", type=int, default=1 << 22)
    a = p.parse_args(argv)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    code = Code()
    h = _lib.Handle(code.H, code.G, device=local)
    for i, e in enumerate(a.ebn0):
        t = run_point(h, e, a.frames, seed=a.seed + i, osd_order=a.order, early_stop=a.early_stop, chunk=a.chunk,
                      max_frame_errors=a.max_errors, rank=rank, world=world)
        if rank == 0:
            print(json.dumps({"ebn0_db": e, **t.as_dict()}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    h.close()


if __name__ == "__main__":
    main()
