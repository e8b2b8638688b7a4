"""Developer probe: what bounds the end-to-end rate with several GPUs on one box -- the PCIe link of each GPU or the host's
ingest bandwidth?  Run under torchrun with N ranks: every rank copies 1 GiB from its GPU into its own pinned host buffer,
first one rank at a time (the link), then all ranks at once (the host).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/d2h_probe_ranks.py
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from pytestbeam_b200.pipeline import bind_host_to_gpu

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
bound = bind_host_to_gpu(local) if "--bind" in sys.argv else False
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, reps = 1 << 30, 6
src = torch.empty(n, dtype=torch.uint8, device="cuda").fill_(3)
dst = torch.empty(n, dtype=torch.uint8).pin_memory()
dst.copy_(src); torch.cuda.synchronize()


def rate():
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    return reps * n / (time.perf_counter() - t) / 1e9


def barrier():
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()

alone = 0.0
for r in range(world):          # one rank at a time: the link
    barrier()
    if r == rank:
        alone = rate()
barrier()
together = rate()               # all at once: the host
barrier()
t = torch.tensor([alone, together], dtype=torch.float64, device="cuda")
parts = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    dist.all_gather(parts, t)
else:
    parts = [t]
if rank == 0:
    a = [float(p[0]) for p in parts]; b = [float(p[1]) for p in parts]
    print(json.dumps({"ranks": world, "numa_bound": bound, "alone_GBps": [round(v, 1) for v in a],
                      "together_GBps": [round(v, 1) for v in b], "sum_together_GBps": round(sum(b), 1),
                      "cpus": os.cpu_count()}), flush=True)
if world > 1:
    dist.destroy_process_group()
