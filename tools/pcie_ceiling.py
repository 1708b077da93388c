"""Platform ceiling for the bench's `e2e` leg: plain pinned-memory copies (no kernels), every rank at once.

    python -m torch.distributed.run --nproc-per-node N tools/pcie_ceiling.py        (or plain python for N = 1)

Per rank: H2D alone, D2H alone, both directions at once (two streams), 2 GiB per direction, best of 3; rank 0 prints
one JSON line with the per-rank and aggregate GB/s.  The bench's e2e moves 14.4 GB up + 7.0 GB down per rank and step."""
import json, os, time
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 1 << 31
h_in = torch.empty(N, dtype=torch.uint8).pin_memory(); h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
d_in = torch.empty(N, dtype=torch.uint8, device="cuda"); d_out = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(mode):
    best = 1e9
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t[0]))
    nbytes = N * (2 if mode == "both" else 1)
    return {"per_rank_gbs": nbytes / best / 1e9, "all_ranks_gbs": world * nbytes / best / 1e9}


res = {m: run(m) for m in ("h2d", "d2h", "both")}
if rank == 0:
    print(json.dumps({"n_gpus": world, "bytes_per_direction": N, "copies": res,
                      "note": "time = max over ranks (all ranks copy at once); pinned host memory via torch (cudaHostAlloc)"}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
