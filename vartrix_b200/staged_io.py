"""Reader for the shard dump the C++ host writes with `vartrix_b200 --dump-staged FILE` (csrc/host/main.cpp)."""
from __future__ import annotations

import struct

import numpy as np

from .engine import StagedBatch

_ORDER = ("locus_row", "hap_bytes", "ref_off", "ref_len", "alt_off", "alt_len", "cand_start", "read_nib", "read_off",
          "read_len", "cb_bytes", "read_cb_off", "read_cb_len", "read_umi_key", "cand_read")
_DT = dict(locus_row=np.uint32, hap_bytes=np.uint8, ref_off=np.uint32, ref_len=np.uint32, alt_off=np.uint32, alt_len=np.uint32,
           cand_start=np.uint64, read_nib=np.uint8, read_off=np.uint64, read_len=np.uint32, cb_bytes=np.uint8,
           read_cb_off=np.uint32, read_cb_len=np.uint16, read_umi_key=np.uint64, cand_read=np.uint32)
METRICS = ("num_reads", "num_low_mapq", "num_non_primary", "num_duplicates", "num_not_useful", "num_invalid_recs",
           "num_multiallelic_recs")


def read_dump(path: str):
    """-> (n_rows, n_cols, [(StagedBatch, host_metrics dict), ...])"""
    data = open(path, "rb").read()
    n_rows, n_cols = struct.unpack_from("<QQ", data, 0)
    p, shards = 16, []
    while p < len(data):
        assert data[p:p + 4] == b"VTXS", "bad shard magic"
        p += 4
        arrs = {}
        for name in _ORDER:
            (nb,) = struct.unpack_from("<Q", data, p); p += 8
            arrs[name] = np.frombuffer(data, dtype=_DT[name], count=nb // np.dtype(_DT[name]).itemsize, offset=p).copy()
            p += nb
        met = dict(zip(METRICS, struct.unpack_from("<7Q", data, p))); p += 56
        shards.append((StagedBatch(n_rows=int(n_rows), **arrs), met))
    return int(n_rows), int(n_cols), shards
