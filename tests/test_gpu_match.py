"""GPU parity, matching stages (K9-K11), through the C ABI: bit-exact against the numpy oracle
(pinned to cv2.BFMatcher in tests/test_oracle_match.py) and against live cv2 when importable."""
import numpy as np
import pytest

from oracle import match as M

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    c = built.Context(640, 240, nfeatures=3000, max_frames=4, ransac_hypotheses=64)
    yield c
    c.close()


def _tie_heavy(rng, n, nbits=10):
    """Descriptors that differ in only a few bits -> many equal distances (tie-break stress)."""
    base = rng.integers(0, 256, 32, dtype=np.uint8)
    d = np.tile(base, (n, 1))
    for i in range(n):
        for b in rng.integers(0, 256, rng.integers(0, nbits)):
            d[i, b // 8] ^= np.uint8(1 << (b % 8))
    return d


@pytest.mark.parametrize("n1,n2,kind", [(1, 1, "rand"), (7, 300, "rand"), (300, 7, "rand"), (257, 513, "ties"),
                                         (1200, 1201, "rand"), (3000, 2999, "ties"), (1025, 64, "ties")])
def test_nn_both_directions_exact(ctx, n1, n2, kind):
    rng = np.random.default_rng(n1 * 7919 + n2)
    if kind == "rand":
        q = rng.integers(0, 256, (n1, 32), dtype=np.uint8); t = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    else:
        q = _tie_heavy(rng, n1); t = _tie_heavy(rng, n2)
    nq, dq, nt, dt = ctx.match_descriptors(q, t)
    onq, odq, ont, odt = M.nearest(q, t)
    assert np.array_equal(nq[0], onq) and np.array_equal(dq[0], odq)
    assert np.array_equal(nt[0], ont) and np.array_equal(dt[0], odt)


def test_batched_pairs_exact(ctx):
    rng = np.random.default_rng(11)
    P, n1, n2 = 5, 700, 650
    q = rng.integers(0, 256, (P, n1, 32), dtype=np.uint8); t = rng.integers(0, 256, (P, n2, 32), dtype=np.uint8)
    nq, dq, nt, dt = ctx.match_descriptors(q, t, npairs=P)
    for p in range(P):
        onq, odq, ont, odt = M.nearest(q[p], t[p])
        assert np.array_equal(nq[p], onq) and np.array_equal(dq[p], odq) and np.array_equal(nt[p], ont) and np.array_equal(dt[p], odt)


def test_empty_sides(ctx):
    q = np.zeros((0, 32), np.uint8); t = np.random.default_rng(0).integers(0, 256, (10, 32), dtype=np.uint8)
    nq, dq, nt, dt = ctx.match_descriptors(q, t)
    assert nq.shape == (1, 0) and (nt == -1).all()


def test_cross_check_matches_cv2_live(ctx):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    q = _tie_heavy(rng, 900, 24); t = _tie_heavy(rng, 950, 24)
    nq, dq, nt, dt = ctx.match_descriptors(q, t)
    i = np.arange(len(q))
    keep = nt[0][nq[0]] == i
    m = cv2.BFMatcher_create(cv2.NORM_HAMMING, True).match(q, t)
    assert np.array_equal(i[keep], [x.queryIdx for x in m])
    assert np.array_equal(nq[0][keep], [x.trainIdx for x in m])
    assert np.array_equal(dq[0][keep], [int(x.distance) for x in m])


@pytest.mark.parametrize("n1,n2", [(2000, 2100), (40, 35), (3000, 3000)])
def test_good_matches_between_slots(ctx, n1, n2):
    """get_good_matches_of_features on identical descriptor matrices (reference visual_odometry.cpp:145-180)."""
    rng = np.random.default_rng(n1)
    t = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    q = t[rng.permutation(n2)[:min(n1, n2)]].copy()
    if n1 > len(q):
        q = np.concatenate([q, rng.integers(0, 256, (n1 - len(q), 32), dtype=np.uint8)])
    flips = rng.integers(0, 60, len(q))           # noisy copies: a spread of distances around the cut
    for i, f in enumerate(flips):
        for b in rng.integers(0, 256, f):
            q[i, b // 8] ^= np.uint8(1 << (b % 8))
    ctx.set_features(0, None, q)
    ctx.set_features(1, None, t)
    good, n_cross, thr = ctx.match(0, 1)
    qi, ti, d, thr_o, nc = M.good_matches(q, t)
    assert n_cross == nc and thr == thr_o
    assert np.array_equal(good["queryIdx"], qi) and np.array_equal(good["trainIdx"], ti)
    assert np.array_equal(good["distance"], d.astype(np.float32)) and (good["imgIdx"] == 0).all()


def test_good_matches_empty_slot(ctx):
    ctx.set_features(0, None, np.zeros((0, 32), np.uint8))
    ctx.set_features(1, None, np.random.default_rng(1).integers(0, 256, (50, 32), dtype=np.uint8))
    good, n_cross, thr = ctx.match(0, 1)
    assert len(good) == 0 and n_cross == 0
    good, n_cross, thr = ctx.match(1, 0)
    assert len(good) == 0 and n_cross == 0


def test_stress_shape_properties(built):
    """Config-4-shaped problem (32k x 32k) at reduced batch: checked through size-independent
    properties -- a planted permutation must be recovered exactly in both directions."""
    import torch
    n = 32768
    g = torch.Generator(device="cuda").manual_seed(1)
    t = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
    perm = torch.randperm(n, device="cuda", generator=g)
    q = t[perm].contiguous()
    q[:, 0] ^= 1                                  # distance exactly 1 to the planted partner
    nnq = torch.empty(n, dtype=torch.int32, device="cuda"); nnt = torch.empty(n, dtype=torch.int32, device="cuda")
    with built.Context(640, 240, nfeatures=500, max_frames=2, ransac_hypotheses=64) as ctx:
        ms = ctx.match_descriptors_device(q.data_ptr(), n, t.data_ptr(), n, 1, nnq.data_ptr(), nnt.data_ptr())
    torch.cuda.synchronize()
    vq = nnq.view(torch.int32).cpu().numpy().view(np.uint32); vt = nnt.cpu().numpy().view(np.uint32)
    assert ms > 0
    assert np.array_equal(vq & 0xFFFF, perm.cpu().numpy()) and (vq >> 16 == 1).all()
    inv = np.empty(n, np.int64); inv[perm.cpu().numpy()] = np.arange(n)
    assert np.array_equal(vt & 0xFFFF, inv) and (vt >> 16 == 1).all()


def test_popcount_variants_agree(built, monkeypatch):
    """Plain 8-POPC, 2-CSA and full Harley-Seal Hamming kernels must return identical neighbours."""
    rng = np.random.default_rng(99)
    q = rng.integers(0, 256, (1500, 32), dtype=np.uint8); t = rng.integers(0, 256, (1400, 32), dtype=np.uint8)
    q[:300] = t[:300]                      # exact duplicates: distance 0 and ties
    ref = M.nearest(q, t)
    for v in ("0", "1", "2"):
        monkeypatch.setenv("B200VO_MATCH_CSA", v)
        with built.Context(640, 240, nfeatures=500, max_frames=2, ransac_hypotheses=64) as c:
            got = c.match_descriptors(q, t)
        for g, r in zip(got, ref):
            assert np.array_equal(g[0], r), "variant " + v
