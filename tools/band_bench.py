"""Throughput of the optional band mode (VTX_BAND_MODEL, csrc/vtx_sw_band.cuh) next to the default full-matrix kernels on the
same device-resident shard (config-3 shape, fewer loci).    python tools/band_bench.py --loci 5000 > profiles/r02_band_bench.json"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--loci", type=int, default=5000)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch
    import vartrix_b200 as vb
    from vartrix_b200 import _capi
    sb, bcs, info = vb.synth.make_shard(n_loci=a.loci, n_barcodes=5000, depth=50, read_len=150, padding=100, seed=3)
    out = dict(what="band mode vs full matrix, consensus, device-resident shard", loci=a.loci, runs=[])
    ref = None
    for name, mode in (("full", _capi.BAND_FULL), ("band_model", _capi.BAND_MODEL)):
        with vb.Engine("consensus", band_mode=mode) as eng:
            eng.set_barcodes(bcs)
            times = []
            for rep in range(a.reps + 1):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                eng.submit(sb)
                res = eng.finish()
                torch.cuda.synchronize()
                times.append(time.perf_counter() - t0)
            pairs = int(res.metrics["num_scored"]) if isinstance(res.metrics, dict) else int(res.metrics.num_scored)
            t = min(times[1:])
            same = None
            if ref is None:
                ref = res
            else:
                import numpy as np
                same = bool(len(ref.row) == len(res.row) and np.array_equal(ref.row, res.row) and np.array_equal(ref.col, res.col) and np.array_equal(ref.val, res.val))
            out["runs"].append(dict(mode=name, pairs=pairs, seconds_host_buffers=t, pairs_per_s=pairs / t, triplets=int(len(res.row)), same_matrix_as_full=same))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
