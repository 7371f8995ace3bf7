"""Host->device copy ceiling of this box for the e2e path: every rank copies the bench's per-step input (2^21 frames x
528 B from pinned memory) with plain cudaMemcpyAsync, all ranks at once, nothing else running.

    python scripts/h2d_ceiling.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/h2d_ceiling.py

Rank 0 prints one JSON line (aggregate GB/s, frames/s equivalent, per-rank GB/s).  bench.py measures the same thing
inside every run and reports it as e2e.h2d_ceiling_GBps / e2e.frac_of_h2d_ceiling."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from short_ldpc_decoding_osd_b200 import _lib
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
numa = _lib.bind_host_to_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
B, steps = 1 << 21, 10
yh, th = _lib.pinned_empty((B, 128), np.float32), _lib.pinned_empty((B, 4), np.int32)
yh[:] = 1.0; th[:] = 0
src, src2 = torch.from_numpy(yh), torch.from_numpy(th)
dst, dst2 = torch.empty((B, 128), dtype=torch.float32, device="cuda"), torch.empty((B, 4), dtype=torch.int32, device="cuda")
def sync():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
for _ in range(2):
    dst.copy_(src, non_blocking=True); dst2.copy_(src2, non_blocking=True)
sync()
t0 = time.perf_counter()
for _ in range(steps):
    dst.copy_(src, non_blocking=True); dst2.copy_(src2, non_blocking=True)
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
per = torch.zeros(world, dtype=torch.float64, device="cuda"); per[rank] = B * 528 * steps / float(dt.item()) / 1e9
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX); dist.all_reduce(per)
if rank == 0:
    agg = world * B * 528 * steps / float(dt.item()) / 1e9
    print(json.dumps({"n_gpus": world, "bytes_per_rank_per_step": B * 528, "steps": steps, "aggregate_GBps": agg, "frames_per_s_equivalent": agg * 1e9 / 528,
                      "per_rank_GBps": [round(x, 2) for x in per.cpu().tolist()], "numa": numa, "host_cpus": len(os.sched_getaffinity(0))}))
if world > 1:
    dist.destroy_process_group()
