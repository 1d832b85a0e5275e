"""GPU (-m gpu): seeded random shapes - ragged corpus sizes, several appends, odd query counts, all scan paths -
against the fp64 arbiter.  Catches tile-edge / chunk-edge / padding mistakes the hand-picked cases may miss."""
import numpy as np
import pytest

from oracle.compare import assert_parity
from oracle.flat_ip import brute_force_fp64

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", list(range(14)))
def test_random_shapes(seed):
    import haconvdr_b200 as hb
    rng = np.random.default_rng(1000 + seed)
    d = int(rng.choice([64, 128, 256, 768, 768, 768]))
    n = int(rng.integers(1, 60000)) if seed % 3 else int(rng.choice([1, 127, 128, 129, 255, 256, 257, 2047, 2048, 2049, 4096]))
    nq = int(rng.integers(1, 400)) if seed % 4 else int(rng.choice([1, 4, 5, 127, 128, 129, 256, 257]))
    k = int(rng.choice([1, 2, 10, 100, 100, 257, 1000]))
    dist = seed % 3
    if dist == 0:
        x = rng.standard_normal((n, d), dtype=np.float32)
        q = rng.standard_normal((nq, d), dtype=np.float32)
    elif dist == 1:                                         # shared mean component
        mu = rng.standard_normal(d).astype(np.float32)
        x = (mu + 0.2 * rng.standard_normal((n, d))).astype(np.float32)
        q = (mu + 0.2 * rng.standard_normal((nq, d))).astype(np.float32)
    else:                                                   # coarse integer grid: many exact ties
        x = rng.integers(-2, 3, size=(n, d)).astype(np.float32)
        q = rng.integers(-2, 3, size=(nq, d)).astype(np.float32)
    idx = hb.FlatIPIndex(d)
    idx.set_option("build_i8", 1)
    cuts = sorted(set([0, n] + [int(c) for c in rng.integers(0, n + 1, size=int(rng.integers(0, 4)))]))
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b > a:
            idx.add(x[a:b])
    assert idx.ntotal == n
    D, I = idx.search(q, k)
    D64, I64 = brute_force_fp64(q, x, k)
    x64 = x.astype(np.float64)
    atol = 1e-5 * float(np.abs(D64[I64 >= 0]).max() + 1e-30) if 4 * k >= n or dist == 2 else 0.0
    if dist == 2:
        assert np.array_equal(I, I64)                      # integer scores: bit-exact ids incl. tie order
    else:
        assert_parity(D64, I64, D, I, rtol=1e-5, atol=atol,
                      ref_scores_of=lambda qi, ids: x64[ids] @ q[qi].astype(np.float64))
    D8, I8 = idx.search(q, k, path=hb.HAC_PATH_I8)
    assert np.array_equal(I8, I) and np.array_equal(D8, D)
    if nq <= 4:
        Dg, Ig = idx.search(q, k, path=hb.HAC_PATH_GEMV)
        assert np.array_equal(Ig, I) and np.array_equal(Dg, D)
