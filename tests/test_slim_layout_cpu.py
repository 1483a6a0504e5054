"""CPU suite: the slim staging layout (vtx_batch2) as the Python host builds it -- the cell-tag code (vtx_pack_cb) is
injective on the tags it accepts and refuses the rest, the vectorised packer agrees with the C function, and slim
shards are self-contained re-packings of the same candidates."""
import numpy as np
import pytest

import vartrix_b200 as vb
from vartrix_b200 import _capi


def test_pack_cb_accepts_exactly_the_documented_form():
    ok = [b"A", b"ACGT", b"ACGTACGTACGTACGT-1", b"T" * 24, b"G" * 24 + b"-99", b"AC-10", b"AC-9"]
    bad = [b"", b"-1", b"ACGN-1", b"acgt", b"A" * 25, b"AC-", b"AC-0", b"AC-01", b"AC-100", b"AC-1x", b"AC_1", b"AC-1 ", b"AC\x00"]
    keys = [vb.pack_cb(s) for s in ok]
    assert all(k < (1 << 60) for k in keys) and len(set(keys)) == len(ok)
    assert all(vb.pack_cb(s) == _capi.NO_CB_KEY for s in bad)
    # injective across lengths and suffixes: "A" vs "AA" vs "A-1", leading A's are not lost
    fam = [b"A", b"AA", b"AAA", b"A-1", b"AA-1", b"A-2", b"A-10", b"C", b"CA", b"AC"]
    assert len({vb.pack_cb(s) for s in fam}) == len(fam)


def test_vectorised_packer_equals_c_function():
    rng = np.random.default_rng(5)
    tags = []
    for i in range(3000):
        n = int(rng.integers(1, 27))
        t = bytes(rng.choice(list(b"ACGT"), n))
        r = rng.random()
        if r < 0.4: t += b"-1"
        elif r < 0.5: t += b"-" + str(int(rng.integers(0, 120))).encode()
        elif r < 0.55: t = t[: n // 2] + b"N" + t[n // 2:]
        elif r < 0.6: t = t.lower()
        tags.append(t)
    tags += [b"ACGTACGTACGTACGT-1"] * 2000          # the common length takes the vectorised path
    off = np.zeros(len(tags), np.uint32); ln = np.zeros(len(tags), np.uint16); pos = 0
    for i, t in enumerate(tags):
        off[i] = pos; ln[i] = len(t); pos += len(t)
    off[7] = 0xFFFFFFFF                              # a read without a tag
    keys, exb, exo = vb.engine.pack_cb_keys(np.frombuffer(b"".join(tags), np.uint8), off, ln)
    n_ex = 0
    for i, t in enumerate(tags):
        if i == 7:
            assert int(keys[i]) == _capi.NO_CB_KEY; continue
        want = vb.pack_cb(t)
        if want == _capi.NO_CB_KEY:
            k = int(keys[i]); assert k & _capi.CB_EXOTIC and k != _capi.NO_CB_KEY
            j = k & 0xFFFFFFFF
            assert exb[exo[j]: exo[j + 1]].tobytes() == t
            n_ex += 1
        else:
            assert int(keys[i]) == want, t
    assert n_ex == len(exo) - 1 and n_ex > 100


@pytest.mark.parametrize("kind,umi", [("snv", False), ("indel", True)])
def test_slim_batch_is_a_repacking(kind, umi):
    sb, bcs, info = vb.synth.make_shard(300, 50, depth=20, seed=9, kind=kind, umi=umi)
    sl = vb.SlimBatch.from_staged(sb, umi)
    assert sl.cand_read is None and (sl.read_umi_key is None) == (not umi)
    assert sl.nbytes() < 0.8 * sb.nbytes()
    units = vb.SlimBatch.units(sl.read_len)
    off = np.concatenate([[0], np.cumsum(units * 4)])
    assert off[-1] == sl.read_nib.size
    for r in (0, 17, sl.n_reads - 1):
        nb = (int(sl.read_len[r]) + 1) // 2
        o = int(sb.read_off[r])
        assert np.array_equal(sl.read_nib[off[r]: off[r] + nb], sb.read_nib[o: o + nb])
    part = sl.shard(40, 90)
    ref = sb.shard(40, 90)
    assert part.n_loci == 50 and part.n_cand == ref.n_cand and part.n_reads == ref.n_reads
    assert np.array_equal(part.read_len, ref.read_len.astype(np.uint16))
    assert np.array_equal(part.cand_start, ref.cand_start) and np.array_equal(part.locus_row, ref.locus_row)
    for l in (0, 49):
        assert np.array_equal(part.hap_bytes[part.ref_off[l]: part.ref_off[l] + part.ref_len[l]],
                              ref.hap_bytes[ref.ref_off[l]: ref.ref_off[l] + ref.ref_len[l]])
        assert np.array_equal(part.hap_bytes[part.alt_off[l]: part.alt_off[l] + part.alt_len[l]],
                              ref.hap_bytes[ref.alt_off[l]: ref.alt_off[l] + ref.alt_len[l]])
