"""GPU: vtx_submit_bam through the ABI (Python mirror), on the host's share of device-staged shards exactly as the CLI hands
them over (`--gpu-stage --dump-staged`): same matrix as the host-staged shards of the same loci, same record-filter counters --
and the error behaviour of the boundary: a corrupt BGZF member, a boundary list that is not made of record boundaries, a
member table that does not tile the stream and a UB string the device cannot key are refused with the documented codes, count
nothing, and leave the context usable."""
import os
import subprocess

import numpy as np
import pytest

from conftest import REF_TEST_DIR, ROOT
from test_host_staging_cpu import _read_vtxd

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "vartrix_b200", "bin", "vartrix_b200")
E_INVALID, E_UNSUPPORTED = -1, -4


def _dumps(tmp_path, pre, bcs, shard, extra=()):
    from vartrix_b200.staged_io import read_dump
    base = [CLI, "-v", f"{REF_TEST_DIR}/{pre}.vcf", "-b", f"{REF_TEST_DIR}/{pre}.bam", "-f", f"{REF_TEST_DIR}/{pre}.fa", "-c", f"{REF_TEST_DIR}/{bcs}",
            "--shard-loci", shard, "--threads", "2", *extra]
    subprocess.run([*base, "--dump-staged", str(tmp_path / "dev.staged"), "--gpu-stage"], check=True, cwd=str(tmp_path))
    subprocess.run([*base, "--dump-staged", str(tmp_path / "host.staged"), "--cut-at-contigs"], check=True, cwd=str(tmp_path))
    n_rows, n_cols, host = read_dump(str(tmp_path / "host.staged"))
    dev = _read_vtxd(str(tmp_path / "dev.staged"))
    assert len(dev) == len(host) and all(d is not None for d in dev)
    return n_rows, dev, host


def _barcodes(bcs):
    import vartrix_b200 as vb
    return vb.Barcodes([ln.strip().encode() for ln in open(f"{REF_TEST_DIR}/{bcs}") if ln.strip()])


@pytest.mark.parametrize("pre,bcs,shard,mode,umi,kw", [
    ("test_dna", "dna_barcodes.tsv", "7", "coverage", False, {}),
    ("test", "barcodes.tsv", "1", "consensus", True, {}),
    ("test_dna", "dna_barcodes.tsv", "1000", "alt_frac", False, dict(mapq=30, primary_only=True, no_duplicates=True)),
])
def test_submit_bam_equals_host_staged_shards(tmp_path, pre, bcs, shard, mode, umi, kw):
    import vartrix_b200 as vb
    extra = (["--umi"] if umi else []) + (["--mapq", "30", "--primary-alignments", "--no-duplicates"] if kw else [])
    n_rows, dev, host = _dumps(tmp_path, pre, bcs, shard, extra)
    b = _barcodes(bcs)
    with vb.Engine(mode, umi=umi) as e_host, vb.Engine(mode, umi=umi) as e_dev:
        e_host.set_barcodes(b); e_dev.set_barcodes(b)
        for d, (hb, _) in zip(dev, host):
            e_host.submit(hb)
            assert e_dev.submit_bam(d, **kw) == 0, e_dev.last_error()
        rh, rd = e_host.finish(), e_dev.finish()
        for f in ("row", "col", "val", "val2", "ref_cnt", "alt_cnt", "unk_cnt"):
            assert np.array_equal(getattr(rh, f), getattr(rd, f), equal_nan=True), f
        assert rh.metrics == rd.metrics and len(rh.row) > 0 and rh.metrics["num_scored"] > 0
        bm = e_dev.bam_metrics()
    for k in ("num_reads", "num_low_mapq", "num_non_primary", "num_duplicates", "num_not_useful"):
        assert bm[k] == sum(int(m[k]) for _, m in host), k


def test_submit_bam_refuses_what_it_must(tmp_path):
    import vartrix_b200 as vb
    _, dev, host = _dumps(tmp_path, "test_dna", "dna_barcodes.tsv", "1000")
    d, (hb, _) = dev[0], host[0]
    assert len(d["members"]) >= 2 and len(d["entry"]) >= 2
    with vb.Engine("coverage") as eng, vb.Engine("coverage") as ref:
        eng.set_barcodes(_barcodes("dna_barcodes.tsv")); ref.set_barcodes(_barcodes("dna_barcodes.tsv"))
        # 1. a flipped bit in the middle of a member's payload: decode error or CRC mismatch
        bad = dict(d); c = bytearray(d["comp"]); m = d["members"][len(d["members"]) // 2]
        c[int(m["in_off"]) + int(m["in_len"]) // 2] ^= 0x10; bad["comp"] = bytes(c)
        assert eng.submit_bam(bad) == E_INVALID and "BGZF member" in eng.last_error()
        # 2. a boundary that is no record boundary: the walk cannot land on it
        bad = dict(d); e = d["entry"].copy(); e[-1] -= 3; bad["entry"] = e
        assert eng.submit_bam(bad) == E_INVALID and "record" in eng.last_error()
        e = np.insert(d["entry"], 1, d["entry"][0] + 5)
        bad = dict(d); bad["entry"] = e
        assert eng.submit_bam(bad) == E_INVALID
        # 3. a member table that does not tile the stream
        bad = dict(d); mm = d["members"].copy(); mm["out_off"][1] += 1; bad["members"] = mm
        assert eng.submit_bam(bad) == E_INVALID and "running sum" in eng.last_error()
        # 4. descending boundaries, loci out of order
        bad = dict(d); bad["entry"] = d["entry"][::-1].copy()
        assert eng.submit_bam(bad) == E_INVALID
        bad = dict(d); bad["row"] = d["row"][::-1].copy()
        assert eng.submit_bam(bad) == E_INVALID
        # nothing of the refused shards was counted, and the context still works: the good shard now equals the host path
        assert eng.bam_metrics()["num_reads"] == 0
        assert eng.submit_bam(d) == 0, eng.last_error()
        ref.submit(hb)
        a, b = eng.finish(), ref.finish()
        assert np.array_equal(a.row, b.row) and np.array_equal(a.col, b.col) and np.array_equal(a.val, b.val) and np.array_equal(a.val2, b.val2)
        assert a.metrics == b.metrics


def test_submit_bam_declines_a_ub_it_cannot_key(tmp_path):
    """with --umi a UB string outside vtx_pack_umi's alphabet sends the shard back (VTX_E_UNSUPPORTED, nothing counted); without
    --umi the same shard is taken"""
    import re
    import zlib
    import vartrix_b200 as vb
    _, dev, host = _dumps(tmp_path, "test", "barcodes.tsv", "1", ["--umi"])
    none = np.uint64(0xFFFFFFFFFFFFFFFF)
    k = next(i for i, (hb, _) in enumerate(host) if len(hb.cand_read) and (hb.read_umi_key[hb.cand_read] != none).any())
    d = dev[k]
    # every UB of the shard's records gets a lower-case first base (same length: offsets, boundaries and ISIZE stay valid)
    comp = bytearray(); mm = d["members"].copy(); n_changed = 0
    for i, m in enumerate(d["members"]):
        raw = zlib.decompress(d["comp"][int(m["in_off"]): int(m["in_off"]) + int(m["in_len"])], -15)
        raw2, n = re.subn(rb"UBZ([ACGTN])([ACGTN]{5,17})\x00", lambda g: b"UBZ" + g.group(1).lower() + g.group(2) + b"\x00", raw)
        n_changed += n
        co = zlib.compressobj(6, zlib.DEFLATED, -15); payload = co.compress(raw2) + co.flush()
        while len(comp) & 7: comp.append(0)
        mm["in_off"][i] = len(comp); mm["in_len"][i] = len(payload); mm["crc"][i] = zlib.crc32(raw2) & 0xFFFFFFFF
        comp += payload
    assert n_changed > 0
    bad = dict(d); bad["members"] = mm; bad["comp"] = bytes(comp) + b"\0" * 16
    b = _barcodes("barcodes.tsv")
    with vb.Engine("consensus", umi=True) as eng:
        eng.set_barcodes(b)
        assert eng.submit_bam(bad) == E_UNSUPPORTED and "UB" in eng.last_error()
        assert eng.bam_metrics()["num_reads"] == 0
        assert eng.submit_bam(d) == 0, eng.last_error()               # the untouched shard right after it
        assert eng.finish().metrics["num_scored"] > 0
    with vb.Engine("consensus", umi=False) as eng:
        eng.set_barcodes(b)
        assert eng.submit_bam(bad) == 0, eng.last_error()
