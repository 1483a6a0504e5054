"""world_size-2 gloo test of the N>1 host logic: per-rank shards, unique-id broadcast, allgatherv assembly.
The compute on each rank is the CPU oracle (no GPU here); the GPU ranks run the same protocol through
vtx_gather (NCCL) -- see bench.py."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import vartrix_b200 as vb
    from vartrix_b200 import dist as vdist
    from oracle import pipeline as P
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base = dict(n_loci=40, n_barcodes=30, depth=15, seed=2, kind="snv", umi=False)
    uid = vdist.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128)
    assert uid == bytes(range(128))
    cfg = vdist.rank_workload(base, rank)
    sb, bcs, info = vb.synth.make_shard(**cfg)
    ob = P.Batch(**{f: getattr(sb, f) for f in P.Batch.FIELDS}, n_rows=sb.n_rows)
    res = P.run_batch(ob, P.Barcodes(bcs.keys), P.MODE_ALT_FRAC, False)
    mine = vb.Triplets(res.row, res.col, res.ref_cnt, res.alt_cnt, res.unk_cnt, res.val, res.val2, res.metrics)
    full = vdist.allgatherv_triplets(mine)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), row=full.row, col=full.col, val=full.val, scored=full.metrics["num_scored"],
             local_rows=res.row, local_scored=res.metrics["num_scored"], bc0=np.frombuffer(bcs.keys[0], np.uint8))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_assemble_row_major(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npz"); r1 = np.load(tmp_path / "rank1.npz")
    for f in ("row", "col"):
        assert np.array_equal(r0[f], r1[f])                                   # every rank holds the whole matrix
    assert np.array_equal(r0["val"], r1["val"], equal_nan=True)
    assert np.array_equal(r0["bc0"], r1["bc0"])                               # same barcode list on every rank
    assert int(r0["scored"]) == int(r0["local_scored"]) + int(r1["local_scored"])
    assert r0["local_rows"].max() < 40 <= r1["local_rows"].min()              # rank r owns rows [40 r, 40 r + 40)
    key = r0["row"].astype(np.int64) * 30 + r0["col"]
    assert (np.diff(key) > 0).all()                                           # rank order == row-major order
    assert len(r0["row"]) == len(r0["local_rows"]) + len(r1["local_rows"])
