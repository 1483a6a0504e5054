"""Informational: wall-clock of the C++ CLI (BAM/VCF/FASTA decode -> GPU -> .mtx) on a synthetic file set.
    python tools/cli_e2e.py --loci 20000 --threads 32
Host decode (BGZF inflate + record parse) bounds this number, not the GPU."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--loci", type=int, default=20000)
    ap.add_argument("--depth", type=int, default=50)
    ap.add_argument("--threads", type=int, nargs="+", default=[8, 32])
    a = ap.parse_args()
    from vartrix_b200 import synth_files
    d = tempfile.mkdtemp(prefix="vtx_cli_")
    t0 = time.time()
    ds = synth_files.write_dataset(d, n_loci=a.loci, n_barcodes=5000, depth=a.depth, read_len=150, seed=3, kind="snv", umi=False, edge_cases=False)
    gen_s = time.time() - t0
    cli = os.path.join(ROOT, "vartrix_b200", "bin", "vartrix_b200")
    out = []
    for th in a.threads:
        for shard in (2048, 8192):
            o = os.path.join(d, f"out_{th}_{shard}.mtx"); r = os.path.join(d, f"ref_{th}_{shard}.mtx")
            t0 = time.time()
            p = subprocess.run([cli, "-v", ds["vcf"], "-b", ds["bam"], "-f", ds["fasta"], "-c", ds["barcodes"], "-o", o, "--ref-matrix", r,
                                "-s", "coverage", "--threads", str(th), "--shard-loci", str(shard), "--log-level", "info"],
                               capture_output=True, text=True)
            dt = time.time() - t0
            scored = [ln for ln in p.stderr.splitlines() if "pairs scored on the GPU" in ln]
            n = int(scored[0].rsplit(" ", 1)[1]) if scored else 0
            phases = [ln.split("] ", 1)[1] for ln in p.stderr.splitlines() if "[INFO] [" in ln]
            out.append(dict(threads=th, shard_loci=shard, wall_s=round(dt, 3), pairs=n, pairs_per_s=round(n / dt) if dt else 0, rc=p.returncode,
                            phases=phases))
    print(json.dumps(dict(reads=ds["n_reads"], loci=a.loci, bam_bytes=os.path.getsize(ds["bam"]), gen_s=round(gen_s, 1), runs=out), indent=1))


if __name__ == "__main__":
    main()
