"""GPU: the optional band mode (vtx_config.band_mode = VTX_BAND_MODEL, csrc/vtx_sw_band.cuh) computes exactly the oracle's
band model (vtxo_sw_band_model: k-mer hits, best chain, +-w band, lazy ends; no hits -> full matrix) -- raw scores pair by
pair and whole matrices -- on the shapes where band and full matrix agree and on those where they do not (short tandem
repeats, indels longer than W).  The default mode stays the full matrix."""
import ctypes
import zlib
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, assert_same_triplets, to_oracle_batch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _staged_from_triples(triples, n_barcodes=7, seed=0):
    """(read, ref_hap, alt_hap) triples, consecutive triples with the same windows form one locus -> StagedBatch + Barcodes"""
    import vartrix_b200 as vb
    rng = np.random.default_rng(seed)
    nibv = {c: v for c, v in zip(b"=ACMGRSVTWYHKDBN", range(16))}
    loci, cur = [], None
    for read, rh, ah in triples:
        if cur is None or cur[0] != rh or cur[1] != ah:
            cur = (rh, ah, []); loci.append(cur)
        cur[2].append(read)
    hap = bytearray(); ref_off, ref_len, alt_off, alt_len, cand_start = [], [], [], [], [0]
    nib = bytearray(); read_off, read_len, cb_off, cb_len, umi, cand = [], [], [], [], [], []
    keys = [(b"ACGTACGTACGT%04d" % i).replace(b"0", b"A").replace(b"1", b"C").replace(b"2", b"G").replace(b"3", b"T")
            .replace(b"4", b"A").replace(b"5", b"C").replace(b"6", b"G").replace(b"7", b"T").replace(b"8", b"A").replace(b"9", b"C") + b"-%d" % (i + 1)
            for i in range(n_barcodes)]
    cbb = bytearray()
    for rh, ah, reads in loci:
        for h, offs, lens in ((rh, ref_off, ref_len), (ah, alt_off, alt_len)):
            while len(hap) & 15: hap.append(0)
            offs.append(len(hap)); lens.append(len(h)); hap += h
        for r in reads:
            while len(nib) & 15: nib.append(0)
            read_off.append(len(nib)); read_len.append(len(r))
            codes = [nibv.get(c, 15) for c in r] + [0]
            nib += bytes((codes[i] << 4) | codes[i + 1] for i in range(0, len(r), 2))
            k = keys[int(rng.integers(0, n_barcodes))]
            cb_off.append(len(cbb)); cb_len.append(len(k)); cbb += k
            umi.append(int(rng.integers(0, 5)) << 5 | 1)
            cand.append(len(read_len) - 1)
        cand_start.append(len(cand))
    while len(hap) & 15: hap.append(0)
    while len(nib) & 15: nib.append(0)
    sb = vb.StagedBatch(np.arange(len(loci)), np.frombuffer(bytes(hap), np.uint8), ref_off, ref_len, alt_off, alt_len, cand_start,
                        np.frombuffer(bytes(nib), np.uint8), read_off, read_len, np.frombuffer(bytes(cbb), np.uint8), cb_off, cb_len, umi, cand,
                        n_rows=len(loci))
    return sb, vb.Barcodes(keys)


def _oracle_band_scores(oracle, triples):
    L = oracle.lib()
    L.vtxo_sw_band_model.restype = ctypes.c_int32
    L.vtxo_sw_band_model.argtypes = [ctypes.c_char_p, ctypes.c_int32, ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
    L.vtxo_sw_full.restype = ctypes.c_int32
    L.vtxo_sw_full.argtypes = [ctypes.c_char_p, ctypes.c_int32, ctypes.c_char_p, ctypes.c_int32]
    band = np.array([[L.vtxo_sw_band_model(r, len(r), rh, len(rh), 6, 20), L.vtxo_sw_band_model(r, len(r), ah, len(ah), 6, 20)] for r, rh, ah in triples], np.int32)
    full = np.array([[L.vtxo_sw_full(r, len(r), rh, len(rh)), L.vtxo_sw_full(r, len(r), ah, len(ah))] for r, rh, ah in triples], np.int32)
    return band, full


FAMILIES = {
    "snv_random": dict(n_loci=12, depth=10, genome="rand"),
    "indel_1_30": dict(n_loci=12, depth=10, genome="rand", indel=(1, 30)),
    "indel_31_60": dict(n_loci=30, depth=10, genome="rand", indel=(31, 60)),
    "snv_str": dict(n_loci=40, depth=10, genome="str"),
    "indel_str": dict(n_loci=40, depth=10, genome="str", indel=(1, 30)),
    "spliced": dict(n_loci=10, depth=10, genome="rand", splice=0.5),
}


@pytest.mark.parametrize("name", list(FAMILIES))
def test_band_mode_scores_equal_the_oracle_band_model(oracle, name):
    import band_exposure as bx
    import vartrix_b200 as vb
    kw = dict(FAMILIES[name]); kw["genome"] = bx.rand_seq if kw["genome"] == "rand" else bx.str_seq
    triples = bx.family(np.random.default_rng(zlib.crc32(name.encode()) % 1000), **kw)        # str hashes change from run to run
    # ragged and tiny reads too: shorter than k (no hits -> full matrix), empty
    triples += [(triples[0][0][:n], triples[0][1], triples[0][2]) for n in (0, 1, 5, 6, 7, 33, 100)]
    sb, bcs = _staged_from_triples(triples)
    band, full = _oracle_band_scores(oracle, triples)
    pair_read = np.arange(sb.n_reads, dtype=np.uint32)
    pair_locus = np.repeat(np.arange(sb.n_loci), np.diff(sb.cand_start).astype(np.int64)).astype(np.uint32)
    with vb.Engine("coverage", band_mode=vb._capi.BAND_MODEL) as eng:
        eng.set_barcodes(bcs)
        rs, as_ = eng.score_pairs(sb, pair_read, pair_locus)
    assert np.array_equal(rs, band[:, 0]) and np.array_equal(as_, band[:, 1])
    with vb.Engine("coverage") as eng:                       # the default mode is the full matrix
        eng.set_barcodes(bcs)
        rs, as_ = eng.score_pairs(sb, pair_read, pair_locus)
    assert np.array_equal(rs, full[:, 0]) and np.array_equal(as_, full[:, 1])
    assert (band <= full).all()
    if name in ("snv_str", "indel_str", "indel_31_60"):
        assert (band < full).any(), "this family is supposed to leave the band"


@pytest.mark.parametrize("mode,umi", [("coverage", False), ("consensus", True), ("alt_frac", False)])
def test_band_mode_matrices_equal_the_oracle(oracle, mode, umi):
    import band_exposure as bx
    import vartrix_b200 as vb
    rng = np.random.default_rng(3)
    triples = bx.family(rng, n_loci=30, depth=12, genome=bx.str_seq) + bx.family(rng, n_loci=15, depth=12, genome=bx.rand_seq, indel=(20, 50))
    sb, bcs = _staged_from_triples(triples, seed=1)
    ob = to_oracle_batch(oracle, sb)
    exp = oracle.run_batch(ob, oracle.Barcodes(bcs.keys), oracle.MODES[mode], umi, n_threads=4, band_model=True)
    full = oracle.run_batch(ob, oracle.Barcodes(bcs.keys), oracle.MODES[mode], umi, n_threads=4)
    with vb.Engine(mode, umi=umi, band_mode=vb._capi.BAND_MODEL) as eng:
        eng.set_barcodes(bcs)
        eng.submit(sb.shard(0, 20)); eng.submit2(vb.SlimBatch.from_staged(sb, umi).shard(20, sb.n_loci))
        got = eng.finish()
    assert_same_triplets(got, exp)
    assert got.metrics == exp.metrics and got.metrics["num_scored"] == full.metrics["num_scored"]


def test_band_mode_constants(oracle):
    import vartrix_b200 as vb
    with pytest.raises(vb.VtxError, match="band constants out of range"):
        vb.Engine("coverage", band_mode=vb._capi.BAND_MODEL, band_k=9)
    with pytest.raises(vb.VtxError, match="unknown band_mode"):
        vb.Engine("coverage", band_mode=7)
