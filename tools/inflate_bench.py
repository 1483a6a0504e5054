"""Throughput of the device BGZF inflate (vtx_bgzf_inflate) on the members of a synthetic BAM, next to zlib on one host
thread: members per call, MB/s of inflated bytes through the synchronous ABI call (host->device copy of the compressed
bytes, kernel, device->host copy of the inflated bytes).
    python tools/inflate_bench.py --loci 20000 > profiles/r02_inflate_bench.json"""
import argparse
import ctypes as C
import json
import os
import struct
import sys
import tempfile
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def members_of(path):
    data = open(path, "rb").read()
    off, out = 0, []
    while off < len(data):
        xlen = struct.unpack_from("<H", data, off + 10)[0]
        total = struct.unpack_from("<H", data, off + 16)[0] + 1
        crc, isize = struct.unpack_from("<II", data, off + total - 8)
        out.append((data[off + 12 + xlen: off + total - 8], isize, crc))
        off += total
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--loci", type=int, default=20000)
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--quals", default="missing", choices=["missing", "binned"], help="binned: i.i.d. four-level qualities, compression ratio of a real BAM")
    a = ap.parse_args()
    import torch
    import vartrix_b200 as vb
    from vartrix_b200 import _capi, synth_files
    d = tempfile.mkdtemp(prefix="vtx_inf_")
    ds = synth_files.write_dataset_fast(d, n_loci=a.loci, n_barcodes=5000, depth=50, level=a.level, quals=a.quals)
    mem = members_of(ds["bam"])
    n = len(mem)
    t0 = time.perf_counter()
    ref = [zlib.decompress(p, -15) for p, _, _ in mem]
    t_zlib = time.perf_counter() - t0
    total_out = sum(len(r) for r in ref); total_in = sum(len(p) for p, _, _ in mem)
    out = dict(what="vtx_bgzf_inflate on the BGZF members of a synthetic BAM (write_dataset_fast)", members=n, compressed_mb=total_in / 1e6, inflated_mb=total_out / 1e6,
               zlib_level=a.level, quals=a.quals, zlib_one_thread_mb_s=total_out / 1e6 / t_zlib, calls=[])
    with vb.Engine("coverage") as eng:
        for per_call in (64, 256, 1024, 4096, n):
            per_call = min(per_call, n)
            # prebuild the call's arrays (pinned) so that only the ABI call is timed
            groups = []
            for g0 in range(0, n, per_call):
                grp = mem[g0:g0 + per_call]
                blocks = (_capi.BgzfBlock * len(grp))()
                comp = bytearray(); out_len = 0
                for i, (payload, isize, crc) in enumerate(grp):
                    while len(comp) & 7: comp.append(0)
                    blocks[i].in_off = len(comp); blocks[i].in_len = len(payload); blocks[i].out_len = isize; blocks[i].out_off = out_len; blocks[i].crc32 = crc
                    comp += payload; out_len += (isize + 15) & ~15
                comp += b"\0" * 16
                ct = torch.frombuffer(comp, dtype=torch.uint8).pin_memory()
                ot = torch.empty(max(out_len, 16), dtype=torch.uint8).pin_memory()
                st = np.zeros(len(grp), np.int32)
                groups.append((blocks, len(grp), ct, len(comp) - 16, ot, out_len, st))
            for crc_flag in (1, 0):
                for rep in range(2):
                    t0 = time.perf_counter()
                    for blocks, k, ct, cl, ot, ol, st in groups:
                        rc = eng._L.vtx_bgzf_inflate(eng._h, blocks, k, ct.data_ptr(), cl, ot.data_ptr(), ol, st.ctypes.data, crc_flag)
                        assert rc == 0, eng._L.vtx_last_error(eng._h)
                    dt = time.perf_counter() - t0
                out["calls"].append(dict(members_per_call=per_call, check_crc=bool(crc_flag), rep=rep, seconds=dt, inflated_mb_s=total_out / 1e6 / dt))
            # correctness of the last group
            blocks, k, ct, cl, ot, ol, st = groups[-1]
            o = ot.numpy()
            g0 = n - k
            for i in (0, k // 2, k - 1):
                assert bytes(o[int(blocks[i].out_off): int(blocks[i].out_off) + int(blocks[i].out_len)]) == ref[g0 + i]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
