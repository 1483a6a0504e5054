"""CPU suite: host I/O edges of the C++ CLI -- the threaded Matrix Market writer against the oracle's writer (Rust `{}`
formatting of f64, main.rs:381-389), and corrupt / truncated BAM input, which must abort the run like the reference's
`let rec = _rec?` (main.rs:830) instead of silently staging fewer reads."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import REF_TEST_DIR, ROOT

CLI = os.path.join(ROOT, "vartrix_b200", "bin", "vartrix_b200")


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hostshim") / "libhost_shim.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "host_shim.cpp"), "-lz", "-lpthread"], check=True)
    lib = ctypes.CDLL(so)
    lib.vtx_test_write_mtx.restype = ctypes.c_int
    lib.vtx_test_write_mtx.argtypes = [ctypes.c_char_p, ctypes.c_ulong, ctypes.c_ulong, ctypes.c_ulong, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_uint]
    return lib


@pytest.mark.parametrize("n,threads", [(0, 1), (5, 4), (70_000, 1), (200_001, 3), (200_001, 16)])
def test_threaded_mtx_writer_equals_oracle_writer(oracle, shim, tmp_path, n, threads):
    rng = np.random.default_rng(n + threads)
    row = np.sort(rng.integers(0, 2**31, n)).astype(np.uint32)
    col = rng.integers(0, 2**32 - 1, n, dtype=np.uint64).astype(np.uint32)
    cnt = rng.integers(0, 40, (n, 2))
    val = np.where(rng.random(n) < 0.5, cnt[:, 0].astype(np.float64), cnt[:, 0] / np.maximum(cnt.sum(1), 1))
    if n > 4:
        val[:5] = [np.nan, 0.0, 1e-7, 123456789012.0, 1.0 / 3.0]
    p = str(tmp_path / "m.mtx")
    assert shim.vtx_test_write_mtx(p.encode(), 1 << 31, (1 << 32) - 1, n, row.ctypes.data, col.ctypes.data, val.ctypes.data, threads) == 1
    assert open(p).read() == oracle.mtx_text(1 << 31, (1 << 32) - 1, row, col, val)


def _stage(tmp_path, bam, *extra):
    out = tmp_path / "d.staged"
    cmd = [CLI, "-v", f"{REF_TEST_DIR}/test.vcf", "-b", str(bam), "-f", f"{REF_TEST_DIR}/test.fa", "-c", f"{REF_TEST_DIR}/barcodes.tsv",
           "--dump-staged", str(out), *extra]
    return subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True)


def test_intact_bam_stages(tmp_path):
    bam = tmp_path / "ok.bam"
    shutil.copy(f"{REF_TEST_DIR}/test.bam", bam); shutil.copy(f"{REF_TEST_DIR}/test.bam.bai", str(bam) + ".bai")
    r = _stage(tmp_path, bam)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("damage", ["flip_payload", "flip_header", "truncate_mid_block", "bad_crc"])
def test_corrupt_bam_aborts_instead_of_staging_less(tmp_path, damage):
    raw = bytearray(open(f"{REF_TEST_DIR}/test.bam", "rb").read())
    # BGZF members: find the second one (the first holds the header) so that the damage lies in fetched record data
    bsize0 = int.from_bytes(raw[16:18], "little") + 1
    if damage == "flip_payload":
        for k in range(bsize0 + 40, bsize0 + 440):
            raw[k] ^= 0x5A
    elif damage == "flip_header":
        raw[bsize0] = 0x00                                    # gzip magic of the second member
    elif damage == "truncate_mid_block":
        raw = raw[: bsize0 + 1000]
    elif damage == "bad_crc":
        bsize1 = int.from_bytes(raw[bsize0 + 16: bsize0 + 18], "little") + 1
        raw[bsize0 + bsize1 - 8] ^= 0xFF                      # CRC32 field of the second member's trailer
    bam = tmp_path / "bad.bam"
    bam.write_bytes(bytes(raw)); shutil.copy(f"{REF_TEST_DIR}/test.bam.bai", str(bam) + ".bai")
    r = _stage(tmp_path, bam)
    assert r.returncode != 0, "corrupt BAM was accepted"
    assert "BGZF" in r.stdout + r.stderr or "BAM" in r.stdout + r.stderr


def test_csi_only_index_is_refused_up_front(tmp_path):
    bam = tmp_path / "c.bam"
    shutil.copy(f"{REF_TEST_DIR}/test.bam", bam); (tmp_path / "c.bam.csi").write_bytes(b"CSI\x01")
    r = _stage(tmp_path, bam)
    assert r.returncode == 1 and "CSI indices are not supported" in r.stderr


def test_gzipped_vcf_is_sniffed_not_named(oracle, tmp_path):
    import gzip
    vcf = tmp_path / "calls.vcf.txt"                          # gzip content, no .gz extension: htslib reads it, so do we
    vcf.write_bytes(gzip.compress(open(f"{REF_TEST_DIR}/test.vcf", "rb").read()))
    out = tmp_path / "d.staged"
    r = subprocess.run([CLI, "-v", str(vcf), "-b", f"{REF_TEST_DIR}/test.bam", "-f", f"{REF_TEST_DIR}/test.fa", "-c", f"{REF_TEST_DIR}/barcodes.tsv",
                        "--dump-staged", str(out)], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    from vartrix_b200.staged_io import read_dump
    n_rows, _, shards = read_dump(str(out))
    assert n_rows == 4 and shards[0][0].n_cand > 0


def test_clmul_crc32_equals_zlib(shim):
    """csrc/host/crc32_fast.hpp (carry-less multiplication, chosen at run time) against zlib.crc32: every length through the
    16- and 64-byte folding boundaries, three alignments, and BGZF-sized buffers"""
    import zlib
    shim.vtx_test_crc32.restype = ctypes.c_uint32
    shim.vtx_test_crc32.argtypes = [ctypes.c_void_p, ctypes.c_ulong]
    rng = np.random.default_rng(3)
    buf = rng.integers(0, 256, size=(1 << 17) + 64, dtype=np.uint8)
    raw = buf.tobytes()
    for off in (0, 1, 7):
        for n in list(range(0, 300)) + [511, 512, 1023, 4096, 65279, 65280, 65535, 65536, 100003]:
            assert shim.vtx_test_crc32(buf.ctypes.data + off, n) == (zlib.crc32(raw[off:off + n]) & 0xFFFFFFFF), (off, n)
    assert shim.vtx_test_crc32(buf.ctypes.data, 0) == 0
