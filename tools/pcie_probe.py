"""pcie_probe.py — host<->device copy ceilings of the box (pinned memory, torch copies on side streams).

    python tools/pcie_probe.py                       # one GPU: one large copy per direction, both directions at once, and
                                                     # 1 GiB cut into 464 / 4096 copies (the per-copy cost the pipelined call pays)
    torchrun --nproc-per-node N tools/pcie_probe.py  # N GPUs at once: every rank copies 1 GiB down and 1 GiB up concurrently,
                                                     # all ranks together (barrier): the host's aggregate ceiling

Numbers of this pool: profiles/r02_e2e_pipeline_trace.txt."""
import os
import time

import torch

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
GB = 1 << 30
n = GB // 8
hd = torch.empty(n, dtype=torch.float64).pin_memory()
hu = torch.empty(n, dtype=torch.float64).pin_memory()
dd = torch.empty(n, dtype=torch.float64, device="cuda")
du = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best


def d2h_one():
    with torch.cuda.stream(s1): hd.copy_(dd, non_blocking=True)


def h2d_one():
    with torch.cuda.stream(s2): du.copy_(hu, non_blocking=True)


def both():
    d2h_one(); h2d_one()


def d2h_pieces(k, streams=1):
    step = n // k
    for i in range(k):
        st = s1 if (streams == 1 or i % 2 == 0) else s3
        with torch.cuda.stream(st): hd[i*step:(i+1)*step].copy_(dd[i*step:(i+1)*step], non_blocking=True)


def d2h_two_halves():
    h = n // 2
    with torch.cuda.stream(s1): hd[:h].copy_(dd[:h], non_blocking=True)
    with torch.cuda.stream(s3): hd[h:].copy_(dd[h:], non_blocking=True)


if world == 1:
    for name, fn, gb in (("D2H 1 GiB one copy", d2h_one, 1), ("H2D 1 GiB one copy", h2d_one, 1), ("D2H + H2D concurrently (1 GiB each)", both, 2),
                         ("D2H 1 GiB as 464 copies, one stream", lambda: d2h_pieces(464), 1), ("D2H 1 GiB as 464 copies, two streams", lambda: d2h_pieces(464, 2), 1),
                         ("D2H 1 GiB as 2 halves on two streams", d2h_two_halves, 1), ("D2H 1 GiB as 4096 copies, one stream", lambda: d2h_pieces(4096), 1)):
        t = timed(fn)
        print(f"{name}: {t*1e3:.2f} ms = {gb * GB / t / 1e9:.1f} GB/s total", flush=True)
else:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def together(fn, reps=5, gb=1):
        """all ranks start together; the slowest rank's time of the best repetition"""
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item()))
        return best

    for name, fn, gb in (("D2H 1 GiB per GPU", d2h_one, 1), ("H2D 1 GiB per GPU", h2d_one, 1), ("D2H + H2D 1 GiB each per GPU", both, 2)):
        t = together(fn)
        if rank == 0:
            print(f"{world} GPUs at once, {name}: {t*1e3:.2f} ms (slowest rank) = {gb * GB / t / 1e9:.1f} GB/s per GPU, "
                  f"{world * gb * GB / t / 1e9:.1f} GB/s aggregate", flush=True)
    try:
        aff = sorted(os.sched_getaffinity(0))
        if rank == 0:
            print(f"host cores visible to a rank: {len(aff)} ({aff[0]}..{aff[-1]})", flush=True)
    except Exception:
        pass
    dist.destroy_process_group()
