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
