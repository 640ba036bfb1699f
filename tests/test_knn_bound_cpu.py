"""CPU property test of the radius estimate k_knn takes from a leaf's distance histogram before it fills its result set
(csrc/photon.cuh, knnSearchWarpT): 256 bins between the nearest and the farthest point of the octant's box, bound = upper edge of the
bin where the running count reaches k. It must be an UPPER bound of the k-th smallest distance of the leaf (so that tightening the
search radius with it cannot drop one of the k nearest photons) - restated here with the kernel's float64 expressions."""
import numpy as np
import pytest

BINS = 256


def bound_from_histogram(d2, lo, hi, k, max_d2=np.inf):
    scale = (BINS - 1) / (hi - lo)
    hist = np.zeros(BINS, dtype=np.int64)
    for v in d2:
        if v <= max_d2:
            q = (v - lo) * scale
            b = 0 if q <= 0.0 else (BINS - 1 if q >= BINS - 1 else int(q))
            hist[b] += 1
    run = np.cumsum(hist)
    if run[-1] < k:
        return None
    bsel = int(np.argmax(run >= k))
    return (lo + (bsel + 1) / scale) * (1.0 + 1e-9)


@pytest.mark.parametrize("seed", range(40))
def test_histogram_bound_is_an_upper_bound_of_the_kth_distance(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(50, 257))
    k = int(rng.integers(1, min(n, 64) + 1))
    bmin = rng.uniform(-5, 5, 3); bmax = bmin + rng.uniform(1e-3, 4, 3)
    style = seed % 4
    if style == 0: pts = rng.uniform(bmin, bmax, (n, 3))
    elif style == 1: pts = bmin + (bmax - bmin) * rng.beta(0.2, 0.2, (n, 3))          # piled up at the faces of the box
    elif style == 2: pts = np.tile(rng.uniform(bmin, bmax, 3), (n, 1))                # coincident photons
    else: pts = np.clip(rng.normal((bmin + bmax) / 2, (bmax - bmin) / 40, (n, 3)), bmin, bmax)
    pts = pts.astype(np.float32).astype(np.float64)                                    # photon positions are stored as float
    bmin, bmax = np.minimum(bmin, pts.min(axis=0)), np.maximum(bmax, pts.max(axis=0))  # octant boxes contain their photons
    p = rng.uniform(bmin - 2, bmax + 2) if seed % 3 else rng.uniform(bmin, bmax)       # query outside / inside the box
    d = p - pts
    d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
    near = np.maximum(np.maximum(bmin - p, p - bmax), 0.0)                             # BoundingBox::distance2 / max_distance2
    far = np.maximum(bmax - p, p - bmin)
    lo, hi = float((near * near).sum()), float((far * far).sum())
    assert hi > lo
    bound = bound_from_histogram(d2, lo, hi, k)
    kth = np.sort(d2)[k - 1]
    assert bound is not None and bound >= kth
    # with a tighter radius already in force only the photons inside it are counted
    cap = float(np.sort(d2)[min(n - 1, k + 5)])
    b2 = bound_from_histogram(d2, lo, hi, k, cap)
    assert b2 is not None and b2 >= kth
    assert bound_from_histogram(d2, lo, hi, k, float(np.sort(d2)[0]) * 0.5 if np.sort(d2)[0] > 0 else -1.0) is None or k == 1
