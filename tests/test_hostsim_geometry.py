"""Host build of the product's geometry header (csrc/geom_core.cuh -> tests/hostsim) checked on
the CPU against the cv2 golden vectors and the numpy oracle.  This is how the five-point solver,
the sampler, the SVD decomposition and the cheirality test are validated without a GPU; the
GPU tests then check the SAME code compiled for sm_100a."""
import ctypes
import os

import numpy as np
import pytest

from oracle import geometry as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hs(built):
    lib = ctypes.CDLL(os.path.join(ROOT, "tests", "hostsim", "libgeom_host.so"))
    return lib


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _pts(golden):
    g = golden["good"]
    K = golden["K"]
    x1 = np.ascontiguousarray(G.normalize_points(golden["f0_pt"][g[:, 0]], K))
    x2 = np.ascontiguousarray(G.normalize_points(golden["f1_pt"][g[:, 1]], K))
    return x1, x2, K


def test_sampler_equals_oracle(hs):
    for (seed, pid, m) in ((42, 0, 60), (1, 999, 5), (2**64 - 1, 12345, 7000)):
        out = np.zeros((32, 5), np.int32)
        hs.hs_draw_samples(ctypes.c_ulonglong(seed), ctypes.c_longlong(pid), 32, m, P(out))
        assert np.array_equal(out, G.draw_samples(seed, pid, 32, m))


def test_real_roots_known_polynomial(hs):
    roots = np.array([-0.9, -0.5, -0.1, 0.2, 0.7])
    c = np.poly(roots)[::-1].copy()                      # ascending
    c = np.concatenate([c, np.zeros(11 - len(c))])
    out = np.zeros(10)
    n = hs.hs_real_roots(P(c), 10, P(out))
    assert n == 5 and np.allclose(out[:5], roots, atol=1e-13)
    # a polynomial with complex pairs only has no real roots
    c2 = np.poly([0.3 + 0.4j, 0.3 - 0.4j, -0.2 + 0.1j, -0.2 - 0.1j])[::-1].real.copy()
    c2 = np.concatenate([c2, np.zeros(11 - len(c2))])
    assert hs.hs_real_roots(P(c2), 10, P(out)) == 0


def test_five_point_vs_cv2_golden(hs, golden):
    x1, x2, K = _pts(golden)
    S = golden["samples"]; ns = golden["fivept_nsol"]; Eg = golden["fivept_E"].reshape(-1, 3, 3)
    off, errs, same, res = 0, [], 0, []
    for h, s in enumerate(S):
        ref = Eg[off:off + ns[h]]; off += ns[h]
        a, b = np.ascontiguousarray(x1[s]), np.ascontiguousarray(x2[s])
        E = np.zeros((10, 9))
        n = hs.hs_five_point(P(a), P(b), P(E))
        got = E[:n].reshape(-1, 3, 3)
        same += n == len(ref)
        for e in ref:
            errs.append(min([min(np.abs(e - g).max(), np.abs(e + g).max()) for g in got], default=1.0))
        for g in got:
            res.append(max(abs(np.linalg.det(g)), np.abs(2 * g @ g.T @ g - np.trace(g @ g.T) * g).max()))
    assert same >= 0.95 * len(S) and (np.array(errs) <= 1e-6).mean() >= 0.90
    assert np.percentile(res, 95) < 1e-12


def test_scoring_decomposition_cheirality_vs_golden(hs, golden):
    x1, x2, K = _pts(golden)
    thr = G.normalized_threshold(K)
    E = np.ascontiguousarray(golden["E_cv"]).reshape(9)
    mask = np.zeros(len(x1), np.uint8)
    c = hs.hs_count_inliers(P(E), P(x1), P(x2), len(x1), ctypes.c_float(np.float32(thr * thr)), P(mask))
    assert c == int(golden["E_mask"].sum()) and np.array_equal(mask, golden["E_mask"])
    R1 = np.zeros(9); R2 = np.zeros(9); t = np.zeros(3)
    hs.hs_decompose(P(E), P(R1), P(R2), P(t))
    best = None
    for (R, tt) in ((R1, t), (R2, t), (R1, -t), (R2, -t)):
        ok = np.zeros(len(x1), np.uint8)
        tt = np.ascontiguousarray(tt)
        hs.hs_cheirality(P(R), P(tt), P(x1), P(x2), len(x1), ctypes.c_double(50.0), P(ok))
        ok &= mask
        if best is None or ok.sum() > best[0]:
            best = (int(ok.sum()), R.reshape(3, 3).copy(), tt.copy(), ok.copy())
    assert best[0] == int(golden["rp_count"]) and np.array_equal(best[3], golden["rp_mask"])
    assert np.abs(best[1] - golden["rp_R"]).max() < 1e-9 and np.abs(best[2] - golden["rp_t"]).max() < 1e-9


def test_cooperative_five_point_vs_cv2_golden(hs, golden):
    """The product's solver (csrc/fivept_group.cuh, 10 lanes per hypothesis) executed on the host with the
    lanes of every phase run one after another, against cv2's solution sets (findEssentialMat on exactly
    five points returns all solutions) and against the independent thread-per-hypothesis form."""
    x1, x2, K = _pts(golden)
    S = golden["samples"]; ns = golden["fivept_nsol"]; Eg = golden["fivept_E"].reshape(-1, 3, 3)
    off, errs, same, res, cross = 0, [], 0, [], []
    for h, s in enumerate(S):
        ref = Eg[off:off + ns[h]]; off += ns[h]
        a, b = np.ascontiguousarray(x1[s]), np.ascontiguousarray(x2[s])
        E = np.zeros((10, 9)); Et = np.zeros((10, 9))
        n = hs.hs_five_point_group(P(a), P(b), P(E))
        nt = hs.hs_five_point(P(a), P(b), P(Et))
        got = E[:n].reshape(-1, 3, 3)
        same += n == len(ref)
        for e in ref:
            errs.append(min([min(np.abs(e - g).max(), np.abs(e + g).max()) for g in got], default=1.0))
        for g in got:
            assert abs(np.linalg.norm(g) - 1.0) < 1e-12
            res.append(max(abs(np.linalg.det(g)), np.abs(2 * g @ g.T @ g - np.trace(g @ g.T) * g).max()))
        for e in Et[:nt]:
            cross.append(min([min(np.abs(e - g.ravel()).max(), np.abs(e + g.ravel()).max()) for g in got], default=1.0))
    assert same >= 0.93 * len(S) and (np.array(errs) <= 1e-6).mean() >= 0.95
    assert np.percentile(res, 95) < 1e-12
    assert (np.array(cross) <= 1e-8).mean() >= 0.97          # the two forms agree on (almost) every root


def test_cooperative_solver_shared_block_layout(hs):
    """2672 bytes: a multiple of 16 (array of blocks stays aligned) and 112 mod 128 (the three groups of a
    warp touch disjoint banks when they read the same member)."""
    n = hs.hs_group_shared_bytes()
    assert n % 16 == 0 and n % 128 == 112 and n <= 4096


def test_dlt_null_vector_by_adjugate_equals_svd(hs):
    """cheirality_ok takes the DLT null vector from three products with adj(G^T G) and CHECKS convergence (the
    Jacobi SVD remains for points that fail it).  Against the oracle's numpy-SVD triangulation: identical
    cheirality masks over scenes from comfortable to no parallax (depth up to 2000 baselines), with pixel noise,
    for all four (R, t) candidates; and the fast path must actually carry the triangulable points."""
    rng = np.random.default_rng(17)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    th = 0.04
    R = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    t = np.array([0.2, -0.1, 1.0]); t /= np.linalg.norm(t)
    fast_total = n_total = 0
    for zmax, noise in ((15, 0.0), (40, 8e-4), (300, 8e-4), (2000, 8e-4), (2000, 0.0)):
        n = 600
        X = np.c_[rng.uniform(-4, 4, n), rng.uniform(-2, 2, n), rng.uniform(2, zmax, n)]
        X[::7, 2] *= -1                                           # some points behind the first camera
        Y = X @ R.T + t
        x1 = np.ascontiguousarray(X[:, :2] / X[:, 2:] + rng.normal(0, noise, (n, 2)))
        x2 = np.ascontiguousarray(Y[:, :2] / Y[:, 2:] + rng.normal(0, noise, (n, 2)))
        tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        R1, R2, tt = G.decompose_essential(tx @ R)              # the four candidates recoverPose votes on
        for Rc, tc in ((R1, tt), (R2, tt), (R1, -tt), (R2, -tt)):
            Rc = np.ascontiguousarray(Rc); tc = np.ascontiguousarray(tc)
            ok = np.zeros(n, np.uint8); fast = np.zeros(n, np.uint8)
            hs.hs_cheirality(P(Rc), P(tc), P(x1), P(x2), n, ctypes.c_double(50.0), P(ok))
            nf = hs.hs_cheirality_fast_path(P(Rc), P(tc), P(x1), P(x2), n, P(fast))
            if zmax <= 40:                                        # triangulable scenes: the fast path must carry them
                fast_total += nf
                n_total += n
            ref = G.cheirality_mask(Rc, tc, x1, x2)
            assert np.array_equal(ok.astype(bool), ref), (zmax, noise, int((ok.astype(bool) != ref).sum()))
    assert fast_total >= 0.98 * n_total, (fast_total, n_total)


def test_register_resident_root_solver_is_bit_identical(hs):
    """The exact root path keeps the level polynomial in registers padded with zero coefficients up to degree 10
    (horner_reg10 / bracket_root_fast_reg10) instead of looping over its true degree from shared memory.  The claim is
    bit-identity, not closeness: compared here on random polynomials of every degree, with roots near the interval
    ends and near-double roots."""
    rng = np.random.default_rng(11)
    hs.hs_bracket_root_both.restype = ctypes.c_int
    solved = 0
    for trial in range(4000):
        deg = int(rng.integers(1, 11))
        if trial % 3 == 0:                                   # a near-double root inside (-1, 1)
            r = rng.uniform(-0.9, 0.9, deg)
            if deg >= 2:
                r[1] = r[0] + rng.choice([1e-3, 1e-6, 1e-9])
            c = np.poly(r)[::-1].copy()
        else:
            c = rng.standard_normal(deg + 1)
        c = np.ascontiguousarray(np.concatenate([c, np.zeros(11 - len(c))]))
        a, b = np.sort(rng.uniform(-1.0, 1.0, 2))
        o1, o2 = ctypes.c_double(), ctypes.c_double()
        rc = hs.hs_bracket_root_both(P(c), deg, ctypes.c_double(a), ctypes.c_double(b), ctypes.byref(o1), ctypes.byref(o2))
        assert rc >= 0, "Horner values differ"
        if rc == 1:
            solved += 1
            assert np.float64(o1.value).tobytes() == np.float64(o2.value).tobytes(), (deg, a, b, o1.value, o2.value)
    assert solved > 1000
