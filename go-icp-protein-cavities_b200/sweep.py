"""Dataset sweep over cavity pairs (replaces the serial one-process-per-pair loop of bo1_GoICP.py:40-54).

Pairs are independent units: rank r of N registers the contiguous block shard_range(n, r, N) on its own GPU with no
data-path collective; only the ~200 B result rows are gathered at the end (torch.distributed, NCCL on GPUs / gloo in
the CPU tests).  One process per GPU.
"""
import numpy as np

ROW = 9 + 3 + 2 + 8   # R, t, (optError, optComp), counters


def shard_range(n, rank, world):
    """contiguous, balanced: the first n % world ranks get one extra pair"""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_results(results):
    out = np.zeros((len(results), ROW), dtype=np.float64)
    for i, r in enumerate(results):
        out[i, :9] = np.asarray(r["R"]).reshape(9)
        out[i, 9:12] = r["t"]
        out[i, 12] = r["optError"]
        out[i, 13] = r["optComp"]
        out[i, 14:22] = r["counters"]
    return out


def unpack_results(rows):
    return [dict(R=row[:9].reshape(3, 3).copy(), t=row[9:12].copy(), optError=float(np.float32(row[12])), optComp=int(row[13]),
                 counters=[int(v) for v in row[14:22]]) for row in rows]


def gather_results(local_rows, n_total, rank, world, device="cpu"):
    """all ranks end up with the (n_total, ROW) table in pair order"""
    if world == 1:
        return local_rows
    import torch
    import torch.distributed as dist
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros((cap, ROW), dtype=torch.float64, device=device)
    buf[: len(local_rows)] = torch.from_numpy(np.ascontiguousarray(local_rows)).to(device)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return np.concatenate([parts[r][: hi - lo].cpu().numpy() for r, (lo, hi) in enumerate(sizes)], axis=0)


def sweep(engine, params, pairs, rank=0, world=1, device="cpu"):
    """registers this rank's block of `pairs` on `engine` and returns the gathered result list for all pairs"""
    lo, hi = shard_range(len(pairs), rank, world)
    local = engine.register_batch(params, pairs[lo:hi]) if hi > lo else []
    rows = gather_results(pack_results(local), len(pairs), rank, world, device)
    return unpack_results(rows)
