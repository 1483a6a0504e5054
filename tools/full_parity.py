"""Full-size parity evidence: the CUDA path against the CPU oracle on complete BASELINE.json-shaped shards
(not just windows of them).  Prints one JSON report; run on the GPU box:
    python tools/full_parity.py config3 config4 config2
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(names):
    import vartrix_b200 as vb
    from oracle import pipeline as P
    threads = len(os.sched_getaffinity(0))
    report = []
    for name in names:
        cfg = dict(vb.synth.CONFIGS[name])
        sb, bcs, info = vb.synth.make_shard(**cfg)
        umi = bool(cfg.get("umi"))
        t0 = time.time()
        with vb.Engine(cfg["scoring_method"], umi=umi) as eng:
            eng.set_barcodes(bcs)
            parts = [sb.shard(lo, hi) for lo, hi in vb.shard_bounds(sb.cand_start, 4)]
            for p in parts:
                eng.submit(p)
            got = eng.finish()
            tiles = eng.tile_counts()
        t_gpu = time.time() - t0
        t0 = time.time()
        ob = P.Batch(**{f: getattr(sb, f) for f in P.Batch.FIELDS}, n_rows=sb.n_rows)
        exp = P.run_batch(ob, P.Barcodes(bcs.keys), P.MODES[cfg["scoring_method"]], umi, n_threads=threads)
        t_cpu = time.time() - t0
        same = all(np.array_equal(getattr(got, f), getattr(exp, f)) for f in ("row", "col", "ref_cnt", "alt_cnt", "unk_cnt")) and \
            np.array_equal(got.val, exp.val, equal_nan=True) and np.array_equal(got.val2, exp.val2, equal_nan=True)
        h = hashlib.sha256()
        for f in ("row", "col", "val"):
            h.update(np.ascontiguousarray(getattr(got, f)).tobytes())
        report.append(dict(config=name, loci=cfg["n_loci"], barcodes=cfg["n_barcodes"], mode=cfg["scoring_method"], umi=umi,
                           pairs=got.metrics["num_scored"], triplets=int(len(got.row)), bit_exact_vs_oracle=bool(same),
                           metrics_equal=bool(got.metrics == exp.metrics), triplets_sha256=h.hexdigest()[:16], sw_tiles_per_kernel_class_last_submit=tiles,
                           gpu_wall_s=round(t_gpu, 2), oracle_wall_s=round(t_cpu, 1), oracle_threads=threads))
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:] or ["config2"])
