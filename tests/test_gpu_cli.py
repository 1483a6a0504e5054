"""GPU end-to-end: the C++ CLI (vartrix flag surface) on real BAM/VCF/FASTA files against the oracle pipeline
run on the same files -- byte-identical Matrix Market text, label files and metrics."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "vartrix_b200", "bin", "vartrix_b200")


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    from vartrix_b200 import synth_files
    d = tmp_path_factory.mktemp("cli_files")
    return synth_files.write_dataset(str(d), n_loci=240, n_barcodes=60, depth=30, read_len=100, seed=11)


@pytest.mark.parametrize("mode,umi,extra,kw", [
    ("consensus", False, [], {}),
    ("alt_frac", False, ["--threads", "3", "--shard-loci", "17"], {}),
    ("coverage", True, ["--threads", "2", "--shard-loci", "50"], {}),
    ("coverage", False, ["--mapq", "20", "--primary-alignments", "--no-duplicates", "--padding", "70"],
     dict(mapq=20, primary_only=True, no_duplicates=True, padding=70)),
    ("consensus", True, ["--bam-tag", "CB", "--valid-chars", "ATGC"], dict(valid_chars="ATGC")),
    ("coverage", True, ["--threads", "3", "--shard-loci", "40", "--gpu-inflate"], {}),        # BGZF members inflated on the device
    ("alt_frac", False, ["--gpu-inflate"], {}),
    # the BAM decoded on the device: inflate, record walk, fetch, record filters, tags (vtx_submit_bam)
    ("coverage", True, ["--threads", "3", "--shard-loci", "40", "--gpu-stage"], {}),
    ("consensus", False, ["--gpu-stage", "--shard-loci", "7", "--threads", "2"], {}),
    ("coverage", False, ["--gpu-stage", "--mapq", "20", "--primary-alignments", "--no-duplicates", "--padding", "70"],
     dict(mapq=20, primary_only=True, no_duplicates=True, padding=70)),
])
def test_cli_matrices_byte_identical_to_oracle(oracle, dataset, tmp_path, mode, umi, extra, kw):
    out, ref, var, bco = (str(tmp_path / n) for n in ("out.mtx", "ref.mtx", "variants.txt", "bcs.tsv"))
    cmd = [CLI, "-v", dataset["vcf"], "-b", dataset["bam"], "-f", dataset["fasta"], "-c", dataset["barcodes"], "-o", out,
           "--ref-matrix", ref, "-s", mode, "--out-variants", var, "--out-barcodes", bco, "--log-level", "info", *extra]
    if umi:
        cmd.append("--umi")
    r = subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    n_rows, n_cols, res, batch, bcs = oracle.run_files(dataset["vcf"], dataset["bam"], dataset["fasta"], dataset["barcodes"],
                                                       mode, umi, n_threads=4, **kw)
    assert open(out).read() == oracle.mtx_text(n_rows, n_cols, res.row, res.col, res.val)
    if mode == "coverage":
        assert open(ref).read() == oracle.mtx_text(n_rows, n_cols, res.row, res.col, res.val2)
    else:
        assert not os.path.exists(ref)                                         # main.rs:385
    recs = oracle.read_vcf(dataset["vcf"])
    assert open(var).read() == "".join(f"{v.chrom}_{v.pos0}\n" for v in recs)   # main.rs:1173-1174 (0-based)
    assert open(bco).read() == "".join(k.decode() + "\n" for k in bcs.keys)      # main.rs:1185-1192
    log = r.stderr
    assert f"Number of alignments evaluated: {batch.host_metrics['num_reads']}" in log
    assert f"not being associated with a cell barcode: {res.metrics['num_not_cell_bc']}" in log
    assert f"not having a UMI: {res.metrics['num_non_umi']}" in log
    assert len(res.row) > 100


def test_gpu_stage_hands_back_what_it_cannot_key(tmp_path):
    """--gpu-stage with --umi on a BAM whose UB strings are partly outside vtx_pack_umi's alphabet (lower case, 20 bases, a dash),
    on two contigs (shards end at the contig change): the device declines the shards that hold such a string, the host stages
    them, and all outputs are byte-identical to the host-staged run."""
    from vartrix_b200.synth_files import BamWriter
    rng = np.random.default_rng(23)
    contigs = [("c1", 30_000), ("c2", 20_000)]
    genome = [rng.integers(0, 4, size=L, dtype=np.uint8) for _, L in contigs]
    acgt = np.frombuffer(b"ACGT", np.uint8)
    fa = tmp_path / "g.fa"
    with open(fa, "wb") as f, open(str(fa) + ".fai", "w") as fai:
        for (name, L), g in zip(contigs, genome):
            f.write(f">{name}\n".encode()); off = f.tell()
            for s0 in range(0, L, 60):
                f.write(acgt[g[s0:s0 + 60]].tobytes() + b"\n")
            fai.write(f"{name}\t{L}\t{off}\t60\t61\n")
    loci = [(ci, int(p)) for ci, (_, L) in enumerate(contigs) for p in np.sort(rng.choice(np.arange(300, L - 300), size=40, replace=False))]
    cbs = [b"ACGTACGTACGTAC" + bytes([b"ACGT"[i & 3], b"ACGT"[(i >> 2) & 3]]) + b"-1" for i in range(12)]
    umis = [b"ACGTACGTAC", b"acgtacgtac", b"ACGTNACGTN", b"ACGTACGTACGTACGTACGT", b"TTTTGGGGCC", b"AC-GT"]      # [1], [3], [5]: not packable
    reads = []
    for ci, p in loci:
        for _ in range(14):
            s0 = p - int(rng.integers(10, 90))
            seq = genome[ci][s0:s0 + 100].copy()
            if rng.random() < 0.5:
                seq[p - s0] = (seq[p - s0] + 1) % 4
            reads.append((ci, s0, acgt[seq].tobytes(), cbs[int(rng.integers(0, len(cbs)))], umis[int(rng.integers(0, len(umis)))]))
    reads.sort(key=lambda r: (r[0], r[1]))
    bam = tmp_path / "r.bam"
    bw = BamWriter(str(bam), contigs)
    for i, (ci, s0, seq, cb, ub) in enumerate(reads):
        bw.add(ci, s0, 60, 0, [("M", 100)], seq, f"q{i}".encode(), b"CBZ" + cb + b"\0" + b"UBZ" + ub + b"\0")
    bw.close()
    vcf = tmp_path / "v.vcf"
    with open(vcf, "w") as f:
        f.write("##fileformat=VCFv4.2\n" + "".join(f"##contig=<ID={n},length={L}>\n" for n, L in contigs) + "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
        for ci, p in loci:
            f.write(f"{contigs[ci][0]}\t{p + 1}\t.\t{'ACGT'[genome[ci][p]]}\t{'ACGT'[(genome[ci][p] + 1) % 4]}\t.\t.\t.\n")
    bcs = tmp_path / "b.tsv"; bcs.write_bytes(b"".join(c + b"\n" for c in cbs))
    outs = {}
    for tag, extra in (("host", []), ("dev", ["--gpu-stage"])):
        d = tmp_path / tag; d.mkdir()
        cmd = [CLI, "-v", str(vcf), "-b", str(bam), "-f", str(fa), "-c", str(bcs), "-o", str(d / "out.mtx"), "--ref-matrix", str(d / "ref.mtx"),
               "-s", "coverage", "--umi", "--shard-loci", "9", "--threads", "2", "--log-level", "info", *extra]
        r = subprocess.run(cmd, cwd=str(d), capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        metric_lines = [ln.split("] ", 1)[-1] for ln in r.stderr.splitlines() if "Number of" in ln]
        outs[tag] = (open(d / "out.mtx").read(), open(d / "ref.mtx").read(), metric_lines, r.stderr)
    assert outs["dev"][0] == outs["host"][0] and outs["dev"][1] == outs["host"][1]
    assert outs["dev"][2] == outs["host"][2] and len(outs["host"][2]) >= 5
    assert "staged on the host after the device declined them" in outs["dev"][3]
    assert outs["host"][0].count("\n") > 100


@pytest.mark.parametrize("extra", [[], ["--gpu-stage"], ["--gpu-inflate"]])
def test_cli_with_a_vcf_without_records(oracle, dataset, tmp_path, extra):
    """main.rs:238-240: zero variants are a warning, the (empty) matrices are still written"""
    vcf = tmp_path / "none.vcf"
    vcf.write_text("##fileformat=VCFv4.2\n##contig=<ID=chr1,length=1000000>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
    out, ref, var = (str(tmp_path / n) for n in ("out.mtx", "ref.mtx", "variants.txt"))
    r = subprocess.run([CLI, "-v", str(vcf), "-b", dataset["bam"], "-f", dataset["fasta"], "-c", dataset["barcodes"], "-o", out, "--ref-matrix", ref,
                        "-s", "coverage", "--out-variants", var, *extra], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Zero variants found in input VCF" in r.stderr
    n_cols = len(oracle.load_barcodes(dataset["barcodes"]))
    empty = oracle.mtx_text(0, n_cols, np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0))
    assert open(out).read() == empty and open(ref).read() == empty and open(var).read() == ""
