"""BASELINE config 5: batch-size sweep of the NMS + order-2 OSD pipeline, 2^10 .. 2^24 frames per call (global),
sharded over the ranks of one node, one 128-byte counter all-reduce per call.

    python scripts/batch_sweep.py                               # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/batch_sweep.py

Each line: global batch, ms per call (CUDA events on the launching stream, max over ranks, median of `reps`
calls after two warm-up calls) and decoded frames/s.  Frames are resident in HBM (Philox generator, Eb/N0 2.5 dB)."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from short_ldpc_decoding_osd_b200 import _lib
from short_ldpc_decoding_osd_b200.fill_matrix_info import Code
from short_ldpc_decoding_osd_b200.simulate import shard_range

ALPHA = 0.66943514
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:  # NCCL's version banner (printf at communicator creation) goes to stderr
    sys.stdout.flush(); _fd = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=torch.device(dev))
    dist.all_reduce(torch.zeros(1, device=dev)); torch.cuda.synchronize()
    sys.stdout.flush(); os.dup2(_fd, 1); os.close(_fd)
code = Code()
h = _lib.Handle(code.H, code.G, device=local)
lo, hi = 10, int(os.environ.get("SWEEP_MAX_LOG2", "24"))
Bmax = (1 << hi) // world + 1
llr = torch.empty((Bmax, 128), dtype=torch.float32, device=dev)
truth = torch.empty((Bmax, 4), dtype=torch.int32, device=dev)
bits = torch.empty((Bmax, 4), dtype=torch.int32, device=dev)
syn = torch.empty((Bmax,), dtype=torch.uint8, device=dev)
cnt = torch.zeros(16, dtype=torch.int64, device=dev)
a0, _ = shard_range(1 << hi, rank, world)
h.call("ldpcb_gen_frames", 7, a0, Bmax, 2.5, llr, truth, None)
stream = torch.cuda.current_stream()
out = []
for lg in range(lo, hi + 1):
    B = 1 << lg
    a, b = shard_range(B, rank, world)
    m = b - a

    def call():
        cnt.zero_()
        if m > 0:
            h.call("ldpcb_decode", llr, m, 12, ALPHA, 1.0, 1.0, 0, 2, 0, bits, syn, None, truth, cnt, stream.cuda_stream)
        if world > 1:
            dist.all_reduce(cnt)

    reps = 20 if lg <= 18 else (8 if lg <= 21 else 4)
    for _ in range(2):
        call()
    ts = []
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); call(); e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([statistics.median(ts)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    frames = int(cnt[0].item())
    assert frames == B, (frames, B)
    if rank == 0:
        rec = {"n_gpus": world, "global_batch": B, "ms_per_call": ms, "frames_per_s": B / (ms * 1e-3)}
        print(json.dumps(rec), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
h.close()
