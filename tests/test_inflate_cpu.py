"""CPU suite: the staging host's own DEFLATE decoder (csrc/host/inflate_fast.hpp) against zlib, differentially.
Every BGZF block of every BAM goes through it (zlib is only the fallback), so it is pinned on random, repetitive,
BAM-like and skewed inputs at every zlib level / strategy, on multi-block streams, and on corrupted / truncated
streams (which it must reject or decode exactly like zlib, never crash)."""
import ctypes
import os
import random
import struct
import subprocess
import zlib

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module", params=["host", "device_logic"])
def inflater(request, tmp_path_factory):
    """host: csrc/host/inflate_fast.hpp (what the staging threads run).  device_logic: the bit-stream half of the device
    decoder csrc/vtx_inflate.cuh (table builders, block headers, symbol batches) compiled for the CPU, with a serial stand-in
    for the warp's apply step (tests/inflate_dev_shim.cpp); the kernel itself is checked on the GPU."""
    d = tmp_path_factory.mktemp("inflate")
    if request.param == "host":
        so = str(d / "libinflate_shim.so")
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "inflate_shim.cpp")], check=True)
        lib = ctypes.CDLL(so)
        lib.vtx_test_inflate.restype = ctypes.c_int
        lib.vtx_test_inflate.argtypes = [ctypes.c_char_p, ctypes.c_ulong, ctypes.c_char_p, ctypes.c_ulong]
        lib.vtx_test_inflate_in_pad.restype = ctypes.c_ulong
        lib.vtx_test_inflate_out_pad.restype = ctypes.c_ulong
        in_pad, out_pad = int(lib.vtx_test_inflate_in_pad()), int(lib.vtx_test_inflate_out_pad())

        def run(comp: bytes, n_out: int):
            inbuf = ctypes.create_string_buffer(comp + b"\xAA" * in_pad, len(comp) + in_pad)
            out = ctypes.create_string_buffer(n_out + out_pad)
            ok = lib.vtx_test_inflate(inbuf, len(comp), out, n_out)
            return bool(ok), out.raw[:n_out]
        return run
    so = str(d / "libinflate_dev_shim.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", cuda_inc, "-o", so, os.path.join(ROOT, "tests", "inflate_dev_shim.cpp")], check=True)
    lib = ctypes.CDLL(so)
    lib.vtx_test_inflate_dev.restype = ctypes.c_int
    lib.vtx_test_inflate_dev.argtypes = [ctypes.c_char_p, ctypes.c_ulong, ctypes.c_char_p, ctypes.c_ulong]

    def run_dev(comp: bytes, n_out: int):
        out = ctypes.create_string_buffer(n_out + 8)
        st = lib.vtx_test_inflate_dev(comp, len(comp), out, n_out)
        return st == 0, out.raw[:n_out]
    return run_dev


def _payload(kind, n, rng, nrng):
    if kind == 0:
        return bytes(nrng.integers(0, 256, n, dtype=np.uint8))                        # incompressible -> stored blocks
    if kind == 1:
        return bytes(nrng.integers(0, 4, n, dtype=np.uint8) + 65)
    if kind == 2:
        return bytes([rng.randrange(256)]) * n                                       # distance-1 runs
    if kind == 3:
        unit = bytes(nrng.integers(0, 256, rng.randrange(1, 12), dtype=np.uint8))    # short-distance overlapping matches
        return (unit * (n // len(unit) + 1))[:n]
    if kind == 4:                                                                     # BAM-like records
        out = bytearray()
        while len(out) < n:
            out += struct.pack("<I", rng.randrange(200, 400)) + bytes(nrng.integers(0, 3, 30, dtype=np.uint8)) + b"read%07d\0" % rng.randrange(10**6)
            out += bytes(nrng.integers(0, 256, 75, dtype=np.uint8)) + bytes(nrng.integers(20, 41, 150, dtype=np.uint8))
            out += b"CBZ" + bytes(nrng.integers(0, 4, 16, dtype=np.uint8) + 65) + b"-1\0"
        return bytes(out[:n])
    p = np.array([2.0 ** -i for i in range(1, 40)]); p /= p.sum()                     # skewed alphabet -> 15-bit codes, second-level tables
    return bytes(nrng.choice(39, size=n, p=p).astype(np.uint8))


def test_inflate_matches_zlib_on_every_block_type(inflater):
    rng = random.Random(1); nrng = np.random.default_rng(1)
    accepted_but_zlib_rejects = 0
    for it in range(1200):
        n = rng.choice([0, 1, 2, 5, 100, 1000, 5000, 20000, 65280, 65535])
        raw = _payload(it % 6, n, rng, nrng)
        level = rng.choice([0, 1, 3, 6, 9])
        strat = rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED])
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strat)
        if rng.random() < 0.2 and n > 10:
            k = rng.randrange(1, n)
            comp = co.compress(raw[:k]) + co.flush(zlib.Z_FULL_FLUSH) + co.compress(raw[k:]) + co.flush()
        else:
            comp = co.compress(raw) + co.flush()
        ok, got = inflater(comp, len(raw))
        assert ok and got == raw, (it, n, level, strat)
        if n:
            assert not inflater(comp, len(raw) - 1)[0]            # the expected size is part of the contract
        assert not inflater(comp, len(raw) + 1)[0]
        if len(comp) > 4:
            for _ in range(2):                                    # single bit flips
                c = bytearray(comp); c[rng.randrange(len(c))] ^= 1 << rng.randrange(8)
                okc, gotc = inflater(bytes(c), len(raw))
                try:
                    ref = zlib.decompress(bytes(c), -15)
                    ref_ok = len(ref) == len(raw)
                except zlib.error:
                    ref_ok = False
                if okc and ref_ok:
                    assert gotc == ref
                accepted_but_zlib_rejects += int(okc and not ref_ok)
            cut = comp[:rng.randrange(len(comp))]
            assert not inflater(cut, len(raw))[0] or len(raw) == 0
    assert accepted_but_zlib_rejects == 0


def test_inflate_decodes_the_reference_bam_fixture_blocks(inflater):
    """Every BGZF member of the committed golden inputs' source format: walk a BGZF file written by the synthetic
    BAM writer and compare each member with zlib."""
    import tempfile
    import vartrix_b200.synth_files as sf
    with tempfile.TemporaryDirectory() as d:
        ds = sf.write_dataset(d, n_loci=60, n_barcodes=20, depth=20, seed=5, edge_cases=True)
        data = open(ds["bam"], "rb").read()
    off, n_blocks = 0, 0
    while off < len(data):
        xlen = struct.unpack_from("<H", data, off + 10)[0]
        total = struct.unpack_from("<H", data, off + 16)[0] + 1
        comp = data[off + 12 + xlen: off + total - 8]
        isize = struct.unpack_from("<I", data, off + total - 4)[0]
        ok, got = inflater(comp, isize)
        assert ok and got == zlib.decompress(comp, -15)
        off += total; n_blocks += 1
    assert n_blocks >= 2
