"""Host-side helpers for the one-process-per-GPU layout (torch.distributed plumbing only).

Loci shard across ranks with no data-path collective; the only exchange is the final
assembly of triplets.  On GPUs that is `vtx_gather` (NCCL allgatherv inside the library).  The same
protocol -- all-gather the counts, then exact-size broadcasts in rank order -- is written here on
torch.distributed tensors so it can be exercised with the gloo backend on CPU-only hosts and used when
the triplets already live on the host."""
from __future__ import annotations

import numpy as np


def rank_workload(base_cfg: dict, rank: int) -> dict:
    """Per-rank synthetic shard of a weak-scaling run: same barcode list, own loci/reads, rows offset by rank."""
    cfg = dict(base_cfg)
    base_seed = cfg["seed"]
    cfg["barcode_seed"] = 1000 + base_seed
    cfg["seed"] = base_seed + 7919 * rank
    cfg["row_offset"] = rank * cfg["n_loci"]
    return cfg


def broadcast_bytes(payload: bytes | None, n: int, src: int = 0, device="cpu") -> bytes:
    """Ship `n` bytes (e.g. the 128-byte NCCL unique id of vtx_comm_unique_id) from rank `src` to every rank."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(n, dtype=torch.uint8)
    if dist.get_rank() == src:
        t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone()
    t = t.to(device)
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


_FIELDS = (("row", np.uint32), ("col", np.uint32), ("ref_cnt", np.uint32), ("alt_cnt", np.uint32), ("unk_cnt", np.uint32),
           ("val", np.float64), ("val2", np.float64))


def allgatherv_triplets(trip, device="cpu"):
    """Allgatherv of per-rank triplets in rank order (= row order for contiguous locus ranges)."""
    import torch
    import torch.distributed as dist
    from .engine import Triplets
    world, rank = dist.get_world_size(), dist.get_rank()
    m = trip.metrics
    mine = torch.tensor([len(trip.row), m.get("num_not_cell_bc", 0), m.get("num_non_umi", 0), m.get("num_scored", 0)],
                        dtype=torch.int64, device=device)
    counts = [torch.zeros(4, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, mine)
    counts = [c.cpu().numpy() for c in counts]
    out = {}
    for name, dt in _FIELDS:
        pieces = []
        for r in range(world):
            n = int(counts[r][0])
            if r == rank:
                buf = torch.from_numpy(np.ascontiguousarray(getattr(trip, name), dtype=dt).view(np.uint8).copy()).to(device)
            else:
                buf = torch.zeros(n * np.dtype(dt).itemsize, dtype=torch.uint8, device=device)
            if n:
                dist.broadcast(buf, r)
            pieces.append(buf.cpu().numpy().view(dt))
        out[name] = np.concatenate(pieces) if pieces else np.zeros(0, dt)
    metrics = dict(num_not_cell_bc=int(sum(c[1] for c in counts)), num_non_umi=int(sum(c[2] for c in counts)),
                   num_scored=int(sum(c[3] for c in counts)))
    return Triplets(metrics=metrics, **out)
