"""Synthetic staged shards of the shapes BASELINE.json names: uniform iid ACGT
context, SNV or 1..30 bp indel loci whose +-padding windows never overlap, `depth` reads per locus
covering the variant (start uniform in [pos-read_len+1, pos]), ref/alt allele 50/50, 0.5 %
substitution errors, cell barcodes uniform over the list with a fraction of unlisted barcodes.

The generator emits the staged `vtx_batch` form directly (what the host stages after BAM/VCF/FASTA
decode and the record filters); numpy's PCG64 with the stated seed makes every shard reproducible.
"""
from __future__ import annotations

import numpy as np

from .engine import Barcodes, StagedBatch

_NIB = np.array([1, 2, 4, 8], np.uint8)            # A C G T in BAM 4-bit code
_ASCII = np.frombuffer(b"ACGT", np.uint8)

CONFIGS = {
    # name: (n_loci, n_barcodes, kind, umi, scoring_method, seed)   -- BASELINE.json configs[1..4]
    "config2": dict(n_loci=10_000, n_barcodes=5_000, kind="snv", umi=False, scoring_method="coverage", seed=1),
    "config3": dict(n_loci=100_000, n_barcodes=50_000, kind="snv", umi=False, scoring_method="consensus", seed=2),
    "config4": dict(n_loci=20_000, n_barcodes=10_000, kind="indel", umi=True, scoring_method="coverage", seed=3),
    "config5_shard": dict(n_loci=62_500, n_barcodes=100_000, kind="snv", umi=False, scoring_method="alt_frac", seed=4,
                          depth=600),
}


def make_barcodes(n_listed: int, n_unlisted: int, rng) -> tuple[Barcodes, np.ndarray]:
    """-> (Barcodes of the listed 16-mers + '-1', uint8 array [n_listed + n_unlisted, 18] of all tags)."""
    need = n_listed + n_unlisted
    vals = np.unique(rng.integers(0, 2**32, size=int(need * 1.3) + 64, dtype=np.uint64))
    vals = rng.permutation(vals)[:need]
    assert len(vals) == need
    shifts = np.arange(15, -1, -1, dtype=np.uint64) * 2
    bases = ((vals[:, None] >> shifts[None, :]) & 3).astype(np.uint8)
    tags = np.empty((need, 18), np.uint8)
    tags[:, :16] = _ASCII[bases]
    tags[:, 16] = ord("-"); tags[:, 17] = ord("1")
    keys = [bytes(tags[i]) for i in range(n_listed)]
    return Barcodes(keys), tags


def make_shard(n_loci: int, n_barcodes: int, depth: int = 50, read_len: int = 150, padding: int = 100, seed: int = 1,
               kind: str = "snv", umi: bool = False, unlisted_frac: float = 0.05, err: float = 0.005,
               row_offset: int = 0, max_indel: int = 30, reads_per_umi: int = 3, n_rows: int | None = None,
               barcode_seed: int | None = None, chunk_loci: int = 4096, **_ignored):
    """-> (StagedBatch, Barcodes, info dict).  Every read is a candidate of exactly one locus."""
    rng = np.random.default_rng(seed)
    brng = np.random.default_rng(seed if barcode_seed is None else barcode_seed)
    n_unlisted = max(16, n_barcodes // 20)
    bcs, tags = make_barcodes(n_barcodes, n_unlisted, brng)

    V = read_len + padding + 8                      # variant column inside the per-locus context
    LCTX = 2 * V + 2 * max_indel + 16
    HS = (2 * padding + 2 + max_indel + 15) // 16 * 16   # haplotype storage stride (16-byte aligned)
    nb = (read_len + 1) // 2
    RS = (nb + 15) // 16 * 16                       # read storage stride
    n_reads = n_loci * depth

    hap = np.zeros((n_loci, 2, HS), np.uint8)
    ref_len = np.zeros(n_loci, np.uint32); alt_len = np.zeros(n_loci, np.uint32)
    read_nib = np.zeros((n_reads, RS), np.uint8)
    cb_idx = np.zeros(n_reads, np.int64)
    umi_key = np.full(n_reads, 0xFFFFFFFFFFFFFFFF, np.uint64)
    n_alt_reads = 0

    for lo in range(0, n_loci, chunk_loci):
        hi = min(lo + chunk_loci, n_loci); n = hi - lo
        ctx = rng.integers(0, 4, size=(n, LCTX), dtype=np.uint8)
        k = np.arange(LCTX + max_indel, dtype=np.int64)[None, :]
        if kind == "snv":
            altb = (ctx[:, V] + rng.integers(1, 4, size=n, dtype=np.uint8)) % 4
            alt_ext = np.concatenate([ctx, ctx[:, :max_indel]], axis=1)
            alt_ext[:, V] = altb
            rl = np.full(n, 2 * padding + 1); al = rl.copy()
            ref_end = np.full(n, V + 1)            # end (exclusive) of the REF allele in ctx coordinates
            alt_end = np.full(n, V + 1)            # end of the ALT allele in alt_ext coordinates
        elif kind == "indel":
            L = rng.integers(1, max_indel + 1, size=n)
            is_ins = rng.random(n) < 0.5
            ins = rng.integers(0, 4, size=(n, max_indel), dtype=np.uint8)
            ctxp = np.concatenate([ctx, rng.integers(0, 4, size=(n, 2 * max_indel), dtype=np.uint8)], axis=1)
            Lc = L[:, None]
            # insertion: ctx[:V+1] + ins[:L] + ctx[V+1:]   deletion: ctx[:V+1] + ctx[V+1+L:]
            src_ins = np.where(k <= V, k, np.where(k <= V + Lc, 0, k - Lc))
            src_del = np.where(k <= V, k, k + Lc)
            src = np.where(is_ins[:, None], src_ins, src_del)
            alt_ext = np.take_along_axis(ctxp, np.minimum(src, ctxp.shape[1] - 1), axis=1)
            ins_zone = is_ins[:, None] & (k > V) & (k <= V + Lc)
            ins_vals = np.take_along_axis(ins, np.clip(k - V - 1, 0, max_indel - 1), axis=1)
            alt_ext = np.where(ins_zone, ins_vals, alt_ext).astype(np.uint8)
            ref_end = np.where(is_ins, V + 1, V + 1 + L)
            alt_end = np.where(is_ins, V + 1 + L, V + 1)
            rl = padding + (ref_end - V) + padding
            al = padding + (alt_end - V) + padding
        else:
            raise ValueError(kind)
        # haplotype windows (construct_haplotypes, main.rs:958-994): [start - pad, end + pad)
        j = np.arange(HS, dtype=np.int64)[None, :]
        ref_idx = np.minimum(V - padding + j, LCTX - 1)
        alt_idx = np.minimum(V - padding + j, alt_ext.shape[1] - 1)
        rh = _ASCII[np.take_along_axis(ctx, ref_idx, axis=1)]
        ah = _ASCII[np.take_along_axis(alt_ext, alt_idx, axis=1)]
        rh[j >= rl[:, None]] = 0; ah[j >= al[:, None]] = 0
        hap[lo:hi, 0] = rh; hap[lo:hi, 1] = ah
        ref_len[lo:hi] = rl; alt_len[lo:hi] = al

        # reads: start uniform in [V - read_len + 1, V]; allele ref/alt 50/50
        nr = n * depth
        loc = np.repeat(np.arange(n), depth)
        start = rng.integers(V - read_len + 1, V + 1, size=nr)
        is_alt = rng.random(nr) < 0.5
        n_alt_reads += int(is_alt.sum())
        idx = start[:, None] + np.arange(read_len, dtype=np.int64)[None, :]
        seq_ref = ctx[loc[:, None], idx]
        seq_alt = alt_ext[loc[:, None], idx]
        seq = np.where(is_alt[:, None], seq_alt, seq_ref).astype(np.uint8)
        n_err = int(rng.binomial(nr * read_len, err))                 # sparse substitution errors
        if n_err:
            ep = rng.integers(0, nr * read_len, size=n_err)
            flat = seq.reshape(-1)
            flat[ep] = (flat[ep] + rng.integers(1, 4, size=n_err, dtype=np.uint8)) % 4
        nib = _NIB[seq]
        if read_len & 1:
            nib = np.concatenate([nib, np.zeros((nr, 1), np.uint8)], axis=1)
        r0 = lo * depth
        read_nib[r0:r0 + nr, :nb] = (nib[:, 0::2] << 4) | nib[:, 1::2]
        listed = rng.random(nr) >= unlisted_frac
        cb_idx[r0:r0 + nr] = np.where(listed, rng.integers(0, n_barcodes, size=nr),
                                      n_barcodes + rng.integers(0, n_unlisted, size=nr))
        if umi:
            n_pool = max(1, -(-depth // reads_per_umi))
            pool = rng.integers(0, 4, size=(n, n_pool, 10), dtype=np.uint64)
            pick = rng.integers(0, n_pool, size=nr)
            ub = pool[loc, pick]                                  # [nr, 10] bases
            sh = (np.arange(9, -1, -1, dtype=np.uint64) * 3)[None, :]
            umi_key[r0:r0 + nr] = ((ub << sh).sum(axis=1).astype(np.uint64) << np.uint64(5)) | np.uint64(10)   # = vtx_pack_umi

    batch = StagedBatch(
        locus_row=(np.arange(n_loci, dtype=np.uint32) + np.uint32(row_offset)),
        hap_bytes=hap.reshape(-1),
        ref_off=(np.arange(n_loci, dtype=np.uint32) * np.uint32(2 * HS)),
        ref_len=ref_len,
        alt_off=(np.arange(n_loci, dtype=np.uint32) * np.uint32(2 * HS) + np.uint32(HS)),
        alt_len=alt_len,
        cand_start=(np.arange(n_loci + 1, dtype=np.uint64) * np.uint64(depth)),
        read_nib=read_nib.reshape(-1),
        read_off=(np.arange(n_reads, dtype=np.uint64) * np.uint64(RS)),
        read_len=np.full(n_reads, read_len, np.uint32),
        cb_bytes=tags[cb_idx].reshape(-1),
        read_cb_off=(np.arange(n_reads, dtype=np.uint32) * np.uint32(18)),
        read_cb_len=np.full(n_reads, 18, np.uint16),
        read_umi_key=umi_key,
        cand_read=np.arange(n_reads, dtype=np.uint32),
        n_rows=(n_rows if n_rows is not None else row_offset + n_loci))
    info = dict(n_loci=n_loci, n_reads=n_reads, n_cand=n_reads, n_pairs=int((cb_idx < n_barcodes).sum()),
                n_alt_reads=n_alt_reads, read_len=read_len, max_hap_len=int(max(ref_len.max(), alt_len.max())),
                kind=kind, depth=depth, seed=seed)
    return batch, bcs, info


def algorithmic_bytes_per_pair(info: dict, padding: int = 100) -> float:
    """ceil(m/2) read nibbles + 16 B pair descriptor + haplotype windows amortised over the
    locus depth + 12 B (4 B scores out + 8 B one 32-bit atomic RMW)."""
    m = info["read_len"]
    n_sum = 2 * (2 * padding + 1) if info["kind"] == "snv" else 2 * (2 * padding + 1) + 16
    return (m + 1) // 2 + 16 + n_sum / info["depth"] + 12
