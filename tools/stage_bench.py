"""Staging-host throughput (no GPU needed): BAM/VCF/FASTA decode + record filters + shard assembly of the C++ CLI,
timed with --dump-staged /dev/null on a synthetic file set, plus the DEFLATE decoder against zlib on the same BAM's
BGZF members.
    python tools/stage_bench.py --loci 20000 --threads 1 2 4 8"""
import argparse
import ctypes
import json
import os
import struct
import subprocess
import sys
import tempfile
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--loci", type=int, default=20000)
    ap.add_argument("--depth", type=int, default=50)
    ap.add_argument("--threads", type=int, nargs="+", default=[1, 2, 4, 8])
    ap.add_argument("--dir", default="")
    a = ap.parse_args()
    from vartrix_b200 import synth_files
    d = a.dir or tempfile.mkdtemp(prefix="vtx_stage_")
    if not os.path.exists(os.path.join(d, "reads.bam")):
        ds = synth_files.write_dataset(d, n_loci=a.loci, n_barcodes=2000, depth=a.depth, seed=7, edge_cases=False)
        n_reads = ds["n_reads"]
    else:
        n_reads = a.loci * a.depth
    cli = os.path.join(ROOT, "vartrix_b200", "bin", "vartrix_b200")
    base = [cli, "-v", f"{d}/variants.vcf", "-b", f"{d}/reads.bam", "-f", f"{d}/genome.fa", "-c", f"{d}/barcodes.tsv",
            "-o", f"{d}/o.mtx", "--dump-staged", "/dev/null", "--log-level", "error"]
    runs = []
    for th in a.threads:
        best = 1e9
        for _ in range(5):
            t0 = time.time()
            subprocess.run(base + ["--threads", str(th)], check=True, capture_output=True)
            best = min(best, time.time() - t0)
        runs.append(dict(threads=th, wall_s=round(best, 3), reads_per_s=round(n_reads / best)))
    # DEFLATE decoder vs zlib on the file's BGZF members
    so = os.path.join(d, "libinflate_shim.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "inflate_shim.cpp")], check=True)
    lib = ctypes.CDLL(so)
    lib.vtx_test_inflate.restype = ctypes.c_int
    lib.vtx_test_inflate.argtypes = [ctypes.c_char_p, ctypes.c_ulong, ctypes.c_char_p, ctypes.c_ulong]
    data = open(f"{d}/reads.bam", "rb").read()
    blocks, off = [], 0
    while off < len(data) and len(blocks) < 600:
        xlen = struct.unpack_from("<H", data, off + 10)[0]
        total = struct.unpack_from("<H", data, off + 16)[0] + 1
        blocks.append((data[off + 12 + xlen: off + total - 8], struct.unpack_from("<I", data, off + total - 4)[0]))
        off += total
    tot = sum(i for _, i in blocks)
    t0 = time.time()
    for _ in range(5):
        for c, _i in blocks:
            zlib.decompress(c, -15)
    tz = (time.time() - t0) / 5
    bufs = [(ctypes.create_string_buffer(c + b"\0" * 16, len(c) + 16), len(c), ctypes.create_string_buffer(i + 64), i) for c, i in blocks]
    t0 = time.time()
    for _ in range(5):
        for ib, n, ob, i in bufs:
            assert lib.vtx_test_inflate(ib, n, ob, i) == 1
    tf = (time.time() - t0) / 5
    print(json.dumps(dict(reads=n_reads, loci=a.loci, bam_bytes=len(data), host_cpus=os.cpu_count(), staging=runs,
                          inflate=dict(blocks=len(blocks), mbytes=round(tot / 1e6, 1), zlib_mb_s=round(tot / tz / 1e6),
                                       own_decoder_mb_s=round(tot / tf / 1e6))), indent=1))


if __name__ == "__main__":
    main()
