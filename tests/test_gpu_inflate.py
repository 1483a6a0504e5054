"""GPU: BGZF members inflated on the device (vtx_bgzf_inflate, csrc/vtx_inflate.cuh: one warp per member, warp-wide match
copies, CRC-32 on the device) against zlib -- the members of the reference's own BAM fixtures, every DEFLATE block type at
every zlib level / strategy, multi-block members, the empty EOF member, and corrupt / truncated members, which must be flagged
member by member without disturbing their neighbours."""
import random
import struct
import zlib

import numpy as np
import pytest

from conftest import REF_TEST_DIR
from test_inflate_cpu import _payload

pytestmark = pytest.mark.gpu


def bgzf_members(path, limit=None):
    data = open(path, "rb").read()
    off, out = 0, []
    while off < len(data) and (limit is None or len(out) < limit):
        xlen = struct.unpack_from("<H", data, off + 10)[0]
        total = struct.unpack_from("<H", data, off + 16)[0] + 1
        payload = data[off + 12 + xlen: off + total - 8]
        crc, isize = struct.unpack_from("<II", data, off + total - 8)
        out.append((payload, isize, crc))
        off += total
    return out


@pytest.fixture(scope="module")
def eng():
    import vartrix_b200 as vb
    with vb.Engine("coverage") as e:
        yield e


@pytest.mark.parametrize("bam,limit", [("test.bam", None), ("test_dna.bam", 400)])
def test_reference_bam_members(eng, bam, limit):
    members = bgzf_members(f"{REF_TEST_DIR}/{bam}", limit)
    got, status = eng.bgzf_inflate(members)
    assert (status == 0).all(), np.nonzero(status)[0][:5]
    for (payload, isize, crc), g in zip(members, got):
        assert len(g) == isize and g == zlib.decompress(payload, -15)
    assert sum(len(g) for g in got) > 100_000 and any(m[1] == 0 for m in members) == (limit is None)   # the EOF member is empty


def test_every_block_type_level_and_strategy(eng):
    rng = random.Random(2); nrng = np.random.default_rng(2)
    members, raws = [], []
    for it in range(600):
        n = rng.choice([0, 1, 2, 5, 100, 1000, 5000, 20000, 65280, 65535, 65536])
        raw = _payload(it % 6, n, rng, nrng)
        co = zlib.compressobj(rng.choice([0, 1, 3, 6, 9]), zlib.DEFLATED, -15, 9,
                              rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]))
        if rng.random() < 0.2 and n > 10:
            k = rng.randrange(1, n)
            comp = co.compress(raw[:k]) + co.flush(zlib.Z_FULL_FLUSH) + co.compress(raw[k:]) + co.flush()
        else:
            comp = co.compress(raw) + co.flush()
        members.append((comp, len(raw), zlib.crc32(raw) & 0xFFFFFFFF)); raws.append(raw)
    got, status = eng.bgzf_inflate(members)
    assert (status == 0).all(), [(int(i), int(status[i])) for i in np.nonzero(status)[0][:5]]
    assert got == raws
    got, status = eng.bgzf_inflate(members, check_crc=False)
    assert (status == 0).all() and got == raws


def test_corrupt_members_are_flagged_one_by_one(eng):
    rng = random.Random(3); nrng = np.random.default_rng(3)
    raws = [_payload(4, 20000 + 1000 * i, rng, nrng) for i in range(40)]          # <= 64 KiB: a BGZF member
    good = [(zlib.compress(r, 6)[2:-4], len(r), zlib.crc32(r) & 0xFFFFFFFF) for r in raws]
    members, expect_bad = [], []
    for i, (comp, n, crc) in enumerate(good):
        kind = i % 5
        if kind == 1:
            c = bytearray(comp); c[rng.randrange(len(c))] ^= 1 << rng.randrange(8); comp = bytes(c)      # bit flip: decode error or CRC mismatch
        elif kind == 2:
            comp = comp[: len(comp) // 2]                                                              # truncated
        elif kind == 3:
            crc ^= 0x1                                                                                 # trailer CRC wrong
        elif kind == 4:
            n -= 1                                                                                     # trailer ISIZE wrong
        members.append((comp, n, crc)); expect_bad.append(kind != 0)
    got, status = eng.bgzf_inflate(members)
    assert [bool(s) for s in status] == expect_bad
    for i, bad in enumerate(expect_bad):
        if not bad:
            assert got[i] == raws[i]
    assert set(int(s) for s in status[np.array(expect_bad)]) <= {1, 2, 3, 4, 5, 6, 7} and 7 in status


def test_bad_descriptors_are_refused(eng):
    import ctypes as C
    import vartrix_b200 as vb
    from vartrix_b200 import _capi
    b = (_capi.BgzfBlock * 1)()
    b[0].in_off = 2; b[0].in_len = 4; b[0].out_len = 4; b[0].out_off = 0
    comp = (C.c_uint8 * 32)(); out = (C.c_uint8 * 8)(); st = (C.c_int32 * 1)()
    rc = eng._L.vtx_bgzf_inflate(eng._h, b, 1, comp, 16, out, 8, st, 1)
    assert rc == -1 and b"4-byte boundary" in eng._L.vtx_last_error(eng._h)
    b[0].in_off = 0; b[0].out_len = 70000
    assert eng._L.vtx_bgzf_inflate(eng._h, b, 1, comp, 16, out, 8, st, 1) == -1
