"""Band exposure report (CPU only): how often would a k=6 / w=20 band (bio 0.30.0's banded aligner as MODELLED by
oracle/vtx_oracle.c::vtxo_sw_band_model -- the crate source is not available here) give a lower score than the full
matrix the engine computes, and how often would that change the call (main.rs:1019-1030)?

    python tools/band_exposure.py > profiles/r02_band_exposure.json

Every family is a set of (read, ref window, alt window) triples built like the reference builds them (padding 100);
reads carry the ref or the alt allele.  Reported per family: pairs, pairs whose banded ref or alt score is below the full
score, the largest deficit, and calls that differ (REF/ALT/UNKNOWN/None)."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pipeline as P   # noqa: E402

L = P.lib()
L.vtxo_sw_band_model.restype = ctypes.c_int32
L.vtxo_sw_band_model.argtypes = [ctypes.c_char_p, ctypes.c_int32, ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
L.vtxo_sw_full.restype = ctypes.c_int32
L.vtxo_sw_full.argtypes = [ctypes.c_char_p, ctypes.c_int32, ctypes.c_char_p, ctypes.c_int32]
L.vtxo_evaluate_scores.restype = ctypes.c_int32
L.vtxo_evaluate_scores.argtypes = [ctypes.c_int32, ctypes.c_int32]
PAD, RL = 100, 150
ACGT = np.frombuffer(b"ACGT", np.uint8)


def rand_seq(rng, n):
    return ACGT[rng.integers(0, 4, n)].tobytes()


def str_seq(rng, n):
    """short tandem repeats: blocks of a 1..6-mer repeated 5..40 times, with random spacers"""
    out = b""
    while len(out) < n:
        unit = rand_seq(rng, int(rng.integers(1, 7)))
        out += unit * int(rng.integers(5, 41)) + rand_seq(rng, int(rng.integers(0, 12)))
    return out[:n]


def mutate(rng, s, err):
    a = np.frombuffer(s, np.uint8).copy()
    hit = np.nonzero(rng.random(len(a)) < err)[0]
    for i in hit:
        a[i] = ACGT[(int(np.searchsorted(ACGT, a[i])) + int(rng.integers(1, 4))) % 4] if a[i] in ACGT else a[i]
    return a.tobytes()


def family(rng, n_loci, depth, genome, indel=(0, 0), err=0.005, splice=0.0, clip=0.0):
    """-> list of (read, ref_hap, alt_hap)"""
    out = []
    for _ in range(n_loci):
        ctx = genome(rng, 2 * (PAD + RL) + 200)
        v = PAD + RL + 20
        lo, hi = indel
        if hi == 0:                                             # SNV
            ref_al, alt_al = ctx[v:v + 1], ACGT[(ACGT.tolist().index(ctx[v]) + 1 + int(rng.integers(0, 3))) % 4: ][:1].tobytes()
        else:
            ln = int(rng.integers(lo, hi + 1))
            if rng.random() < 0.5:                               # insertion (VCF style: anchor base + inserted bases)
                ref_al, alt_al = ctx[v:v + 1], ctx[v:v + 1] + rand_seq(rng, ln)
            else:                                                # deletion
                ref_al, alt_al = ctx[v:v + 1 + ln], ctx[v:v + 1]
        end = v + len(ref_al)
        ref_hap = ctx[v - PAD:end + PAD]
        alt_hap = ctx[v - PAD:v] + alt_al + ctx[end:end + PAD]
        for _ in range(depth):
            src = (ctx[:v] + alt_al + ctx[end:]) if rng.random() < 0.5 else ctx
            start = int(rng.integers(v - RL + 1, v + 1))
            read = src[start:start + RL]
            if rng.random() < splice:                            # the read continues in another exon: its tail is unrelated sequence
                cut = int(rng.integers(RL // 3, RL))
                read = read[:cut] + rand_seq(rng, RL - cut)
            if rng.random() < clip:                              # soft-clipped adapter / poly-A tail
                cut = int(rng.integers(RL // 2, RL))
                read = read[:cut] + b"A" * (RL - cut)
            out.append((mutate(rng, read, err), ref_hap, alt_hap))
    return out


def evaluate(triples):
    n = len(triples); diff = flips = 0; worst = 0
    for read, rh, ah in triples:
        rf, af = L.vtxo_sw_full(read, len(read), rh, len(rh)), L.vtxo_sw_full(read, len(read), ah, len(ah))
        rb, ab = L.vtxo_sw_band_model(read, len(read), rh, len(rh), 6, 20), L.vtxo_sw_band_model(read, len(read), ah, len(ah), 6, 20)
        assert rb <= rf and ab <= af
        if rb != rf or ab != af:
            diff += 1; worst = max(worst, rf - rb, af - ab)
            if L.vtxo_evaluate_scores(rf, af) != L.vtxo_evaluate_scores(rb, ab):
                flips += 1
    return dict(pairs=n, pairs_band_below_full=diff, largest_deficit=worst, calls_that_differ=flips)


def main():
    rng = np.random.default_rng(20)
    fams = {
        "config 2/3 shape: SNVs, random genome, 0.5 % errors": dict(n_loci=300, depth=20, genome=rand_seq),
        "config 4 shape: indels of 1-30 bases (W = 20)": dict(n_loci=300, depth=20, genome=rand_seq, indel=(1, 30)),
        "indels of 31-60 bases": dict(n_loci=200, depth=20, genome=rand_seq, indel=(31, 60)),
        "SNVs in short tandem repeats / low complexity": dict(n_loci=300, depth=20, genome=str_seq),
        "indels of 1-30 bases in short tandem repeats": dict(n_loci=300, depth=20, genome=str_seq, indel=(1, 30)),
        "SNVs, spliced reads (tail continues in another exon)": dict(n_loci=200, depth=20, genome=rand_seq, splice=0.5),
        "SNVs, soft-clipped poly-A tails": dict(n_loci=200, depth=20, genome=rand_seq, clip=0.5),
        "SNVs, 5 % substitution errors": dict(n_loci=200, depth=20, genome=rand_seq, err=0.05),
    }
    report = {"what": "vtxo_sw_band_model (k=6, w=20; a MODEL of bio 0.30.0's band, golden-consistent, crate source unavailable) vs the full matrix "
                      "the engine computes; banded <= full always",
              "families": {}}
    for name, kw in fams.items():
        report["families"][name] = evaluate(family(rng, **kw))
        print(name, report["families"][name], file=sys.stderr)
    tot = {k: sum(f[k] for f in report["families"].values()) for k in ("pairs", "pairs_band_below_full", "calls_that_differ")}
    report["total"] = tot
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
