"""CPU suite: the decomposition the folded GPU kernel (vtx_sw_fold.cuh) relies on.

score(read, hap) = max( forward DP over hap[:P], DP of the reversed read over reversed hap[n-S:],
                        forward DP continued over hap[P:n-S], junction terms )        for ANY P + S <= n.
oracle.sw_fold restates that on the CPU; here it is pinned against the full-matrix oracle."""
import numpy as np
import pytest

ACGT = np.frombuffer(b"ACGT", np.uint8)


def _mutated(rng, src, sub=0.02, dele=0.01, ins=0.01):
    out = []
    for b in src:
        u = rng.random()
        if u < sub:
            out.append(ACGT[rng.integers(0, 4)])
        elif u < sub + dele:
            continue
        elif u < sub + dele + ins:
            out.append(b); out.extend(ACGT[rng.integers(0, 4, int(rng.integers(1, 8)))])
        else:
            out.append(b)
    return np.array(out if out else [ACGT[0]], np.uint8)


@pytest.mark.parametrize("alphabet", [b"ACGT", b"AC", b"A"])
def test_fold_decomposition_equals_full_matrix(oracle, alphabet):
    rng = np.random.default_rng(len(alphabet))
    alpha = np.frombuffer(alphabet, np.uint8)
    checked = 0
    for it in range(1500):
        n = int(rng.integers(1, 260))
        hap = alpha[rng.integers(0, len(alpha), n)]
        if it % 5 == 0:
            read = alpha[rng.integers(0, len(alpha), int(rng.integers(1, 170)))]
        else:
            st = int(rng.integers(-30, max(1, n - 10))); m = int(rng.integers(1, 170))
            src = [hap[j] if 0 <= j < n else alpha[rng.integers(0, len(alpha))] for j in range(st, st + m)]
            read = _mutated(rng, src)
            if it % 5 == 4 and len(read) > 20:          # a long deletion somewhere
                c = int(rng.integers(5, len(read) - 5)); read = np.concatenate([read[:c], read[c + int(rng.integers(1, 25)):]])
        x, y = read.tobytes(), hap.tobytes()
        full = oracle.sw_full(x, y)
        for _ in range(3):
            p = int(rng.integers(0, n + 1)); s = int(rng.integers(0, n - p + 1))
            assert oracle.sw_fold(x, y, p, s) == full, (it, len(x), n, p, s)
            checked += 1
        # the split the kernel uses: both flanks 96 columns
        if n >= 193:
            assert oracle.sw_fold(x, y, 96, 96) == full
    assert checked == 4500


def test_fold_rejects_overlapping_flanks(oracle):
    assert oracle.sw_fold(b"ACGT", b"ACGTACGT", 5, 4) == -1
