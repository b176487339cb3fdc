// TEST INFRASTRUCTURE ONLY -- never included by the product.
// The thread-per-hypothesis form of the five-point solver (round-1 v1/v2 of the CUDA path, replaced by
// the cooperative solver in csrc/fivept_group.cuh): null space by Gaussian elimination + Gram-Schmidt,
// 10x20 Gauss-Jordan in a per-thread array, depth-first Sturm isolation with the exact derivative-chain
// fallback.  Kept here as an independent second restatement: tests/test_hostsim_geometry.py checks
// both forms against the cv2 golden vectors and against each other.
#pragma once
#include "../../slam-visual-odometry_b200/csrc/geom_core.cuh"

namespace b200vo {
namespace geom {

// root of c (degree deg) inside the sign-changing bracket (a,b): Newton, bisection when unsafe
GEOM_HD inline double bracket_root(const double* c, int deg, double a, double b, double fa) {
    double x = 0.5 * (a + b);
    for (int it = 0; it < 80; ++it) {
        double f, df;
        horner2(c, deg, x, f, df);
        if (f == 0.0) return x;
        if ((f < 0.0) == (fa < 0.0)) a = x; else b = x;
        double xn = x - f / df;
        if (!(xn > a && xn < b)) xn = 0.5 * (a + b);
        if (fabs(xn - x) <= 4e-16 * fabs(x) + 1e-300 || b - a <= 2e-16 * fabs(x)) return xn;
        x = xn;
    }
    return x;
}

// All simple real roots of c[0..deg] inside [-1, 1], ascending.  MAXDEG >= deg.
template <int MAXDEG>
GEOM_HD inline int real_roots_unit(const double* c, int deg, double* roots) {
    while (deg > 0 && c[deg] == 0.0) --deg;
    if (deg <= 0) return 0;
    double prev[MAXDEG], cur[MAXDEG], d[MAXDEG + 1];
    int nprev = 0;
    // derivative chain from the (deg-1)-th derivative (linear) up to the polynomial itself;
    // the k-th derivative scaled by 1/k! has coefficients c[i+k] * C(i+k, k)
    for (int k = deg - 1; k >= 0; --k) {
        const int dk = deg - k;
        for (int i = 0; i <= dk; ++i) {
            double binom = 1.0;
            for (int j = 1; j <= k; ++j) binom = binom * (double)(i + j) / (double)j;
            d[i] = c[i + k] * binom;
        }
        int ncur = 0;
        double lo = -1.0, flo = horner(d, dk, lo);
        for (int s = 0; s <= nprev; ++s) {
            const double hi = (s < nprev) ? prev[s] : 1.0;
            const double fhi = horner(d, dk, hi);
            if (flo == 0.0) {
                if (s == 0) cur[ncur++] = lo;            // root exactly at -1
            } else if (fhi == 0.0) {
                if (ncur < MAXDEG) cur[ncur++] = hi;
            } else if ((flo < 0.0) != (fhi < 0.0)) {
                if (ncur < MAXDEG) cur[ncur++] = bracket_root_fast(d, dk, lo, hi, flo);
            }
            lo = hi;
            flo = fhi;
        }
        nprev = ncur;
        for (int i = 0; i < ncur; ++i) prev[i] = cur[i];
    }
    for (int i = 0; i < nprev; ++i) roots[i] = prev[i];
    return nprev;
}

// Sturm-sequence isolation of the simple real roots of c[0..deg] in (-1, 1], ascending.
// p0 = p, p1 = p', p_{k-1} = (a_k x + b_k) p_k - p_{k+1}; every remainder is rescaled by a positive
// factor s_k (sign counts are unaffected).  The chain is evaluated by its three-term recurrence
// (46 FMAs per point for degree 10).  Returns -1 when a degree drops by more than one (leading
// coefficient of a remainder ~ 0): the caller then falls back to real_roots_unit.
template <int MAXDEG>
struct SturmChain {
    double a[MAXDEG + 1], b[MAXDEG + 1], s[MAXDEG + 1];
    double dp[MAXDEG + 1];           // p'
    const double* p;
    int n;                            // degree of p; chain has n+1 members

    GEOM_HD int build(const double* c, int deg) {
        p = c; n = deg;
        double f[MAXDEG + 1], g[MAXDEG + 1];
        for (int i = 0; i <= deg; ++i) f[i] = c[i];
        for (int i = 0; i < deg; ++i) { dp[i] = c[i + 1] * (double)(i + 1); g[i] = dp[i]; }
        int m = deg;                  // deg f = m, deg g = m-1
        for (int k = 1; k < deg; ++k, --m) {
            const double lead = g[m - 1];
            double gmax = 0.0;
            for (int i = 0; i < m; ++i) gmax = fmax(gmax, fabs(g[i]));
            if (!(fabs(lead) > 1e-13 * gmax) || !(gmax > 0.0)) return -1;
            const double inv = 1.0 / lead;
            const double ak = f[m] * inv;
            const double bk = (f[m - 1] - ak * ((m >= 2) ? g[m - 2] : 0.0)) * inv;
            // r = f - (ak x + bk) g  (degree <= m-2);  p_{k+1} = -r * s
            double rmax = 0.0;
            for (int i = 0; i <= m - 2; ++i) {
                const double gi1 = (i >= 1) ? g[i - 1] : 0.0;
                const double r = f[i] - ak * gi1 - bk * g[i];
                f[i] = -r;                                     // reuse f as the new remainder
                rmax = fmax(rmax, fabs(r));
            }
            if (!(rmax > 0.0)) return -1;                      // exact common factor (multiple root)
            const double sk = 1.0 / rmax;
            a[k] = ak; b[k] = bk; s[k] = sk;
            // rotate: f <- g, g <- scaled remainder
            for (int i = 0; i <= m - 2; ++i) { const double t = f[i] * sk; f[i] = g[i]; g[i] = t; }
            f[m - 1] = g[m - 1];      // g had m coefficients (0..m-1); slot m-1 was not overwritten above
        }
        return 0;
    }

    // number of sign changes of the chain at x
    GEOM_HD int changes(double x) const {
        double v0 = p[n], v1 = dp[n - 1];
        for (int i = n - 1; i >= 0; --i) v0 = v0 * x + p[i];
        for (int i = n - 2; i >= 0; --i) v1 = v1 * x + dp[i];
        int cnt = 0;
        int last = (v0 > 0.0) - (v0 < 0.0);
        int sg = (v1 > 0.0) - (v1 < 0.0);
        if (sg != 0) { if (last != 0 && sg != last) ++cnt; last = sg; }
        for (int k = 1; k < n; ++k) {
            const double v2 = s[k] * ((a[k] * x + b[k]) * v1 - v0);
            v0 = v1; v1 = v2;
            sg = (v2 > 0.0) - (v2 < 0.0);
            if (sg != 0) { if (last != 0 && sg != last) ++cnt; last = sg; }
        }
        return cnt;
    }
};

// Brackets (lo, hi] that each contain exactly one simple real root of c in (-1, 1], ascending.
// lo == hi marks a root that is already known exactly (end point, unresolved cluster, or the
// derivative-chain fallback, which returns refined roots).
template <int MAXDEG>
GEOM_HD inline int real_root_brackets(const double* c, int deg, double* lo, double* hi, bool allow_fallback = true) {
    while (deg > 0 && c[deg] == 0.0) --deg;
    if (deg <= 0) return 0;
    if (deg == 1) { const double r = -c[0] / c[1]; if (r > -1.0 && r <= 1.0) { lo[0] = hi[0] = r; return 1; } return 0; }
    SturmChain<MAXDEG> ch;
    bool ok = (ch.build(c, deg) == 0);
    double sa[26], sb[26];
    signed char va[26], vb[26];
    int top = 0, nr = 0;
    if (ok) {
        sa[0] = -1.0; sb[0] = 1.0; va[0] = (signed char)ch.changes(-1.0); vb[0] = (signed char)ch.changes(1.0);
        // parity check: the root count in (-1,1] is odd iff p changes sign between the ends.  A
        // numerically unreliable chain (near-multiple roots: last members ~ 0) fails it -> fallback.
        const double fm = horner(c, deg, -1.0), fp = horner(c, deg, 1.0);
        const int k = va[0] - vb[0];
        const bool odd = (fm < 0.0) != (fp < 0.0);
        if (k < 0 || fm == 0.0 || fp == 0.0 || ((k & 1) != 0) != odd) ok = false;
        top = 1;
    }
    if (!ok && !allow_fallback) return -1;      // the caller re-queues this polynomial for the slow path
    if (!ok) {
        // Unreliable chain (near-multiple roots): the derivative-chain method is exact but slow and
        // divergent.  (A fixed-grid sign scan was tried as a warp-uniform fallback: it loses the
        // close root pairs in up to 8 % of the hypotheses of a small scene -- rejected.)
        double r[MAXDEG];
        const int n = real_roots_unit<MAXDEG>(c, deg, r);
        for (int i = 0; i < n; ++i) lo[i] = hi[i] = r[i];
        return n;
    }
    // depth-first interval splitting; left interval first so the roots come out ascending
    while (top > 0 && nr < MAXDEG) {
        --top;
        const double a = sa[top], b = sb[top];
        const int ca = va[top], cb = vb[top];
        const int k = ca - cb;
        if (k <= 0) continue;
        if (k == 1) {
            const double fa = horner(c, deg, a), fb = horner(c, deg, b);
            if (fb == 0.0) { lo[nr] = hi[nr] = b; ++nr; continue; }
            if (fa != 0.0 && (fa < 0.0) != (fb < 0.0)) { lo[nr] = a; hi[nr] = b; ++nr; continue; }
            // fa == 0 (root sits on the open end) or inconsistent signs: split further
        }
        if (b - a < 5e-7 || top + 2 > 26) { lo[nr] = hi[nr] = 0.5 * (a + b); ++nr; continue; }   // unresolved cluster: report once
        const double m = 0.5 * (a + b);
        const int cm = ch.changes(m);
        sa[top] = m; sb[top] = b; va[top] = (signed char)cm; vb[top] = (signed char)cb; ++top;        // right half (popped second)
        sa[top] = a; sb[top] = m; va[top] = (signed char)ca; vb[top] = (signed char)cm; ++top;        // left half (popped first)
    }
    return nr;
}

template <int MAXDEG>
GEOM_HD inline int real_roots_sturm(const double* c, int deg, double* roots) {
    double lo[MAXDEG], hi[MAXDEG];
    const int n = real_root_brackets<MAXDEG>(c, deg, lo, hi);
    for (int i = 0; i < n; ++i)
        roots[i] = (lo[i] == hi[i]) ? lo[i] : bracket_root_fast(c, deg, lo[i], hi[i], horner(c, deg, lo[i]));
    return n;
}

// Storage policies for the 10x20 constraint matrix.  The product keeps it in SHARED memory,
// element-major / thread-minor (conflict-free, and the only way the working set stays on chip:
// 1600 B per hypothesis); the host test build uses a plain array.
struct LocalMat {
    double m[200];
    GEOM_HD double& operator()(int i, int j) { return m[i * 20 + j]; }
};
struct StridedMat {
    double* base;      // &smem[thread]
    int stride;        // threads per CTA
    GEOM_HD double& operator()(int i, int j) { return base[(i * 20 + j) * stride]; }
};

// Orthonormal basis (4 x 9) of the null space of the 5x9 epipolar system.
GEOM_HD inline void null_space_5x9(double A[5][9], double N[4][9]) {
    int pc[5];
    bool used[9];
    for (int j = 0; j < 9; ++j) used[j] = false;
    for (int r = 0; r < 5; ++r) {
        // complete pivoting over the remaining rows and unused columns
        int pr = r, pcn = -1;
        double best = -1.0;
        for (int i = r; i < 5; ++i)
            for (int j = 0; j < 9; ++j)
                if (!used[j] && fabs(A[i][j]) > best) { best = fabs(A[i][j]); pr = i; pcn = j; }
        if (pr != r)
            for (int j = 0; j < 9; ++j) { const double t = A[r][j]; A[r][j] = A[pr][j]; A[pr][j] = t; }
        used[pcn] = true;
        pc[r] = pcn;
        const double inv = (best > 0.0) ? 1.0 / A[r][pcn] : 0.0;
        for (int j = 0; j < 9; ++j) A[r][j] *= inv;
        for (int i = 0; i < 5; ++i) {
            if (i == r) continue;
            const double f = A[i][pcn];
            if (f != 0.0)
                for (int j = 0; j < 9; ++j) A[i][j] -= f * A[r][j];
        }
    }
    int k = 0;
    for (int f = 0; f < 9; ++f) {
        if (used[f]) continue;
        for (int j = 0; j < 9; ++j) N[k][j] = 0.0;
        N[k][f] = 1.0;
        for (int r = 0; r < 5; ++r) N[k][pc[r]] = -A[r][f];
        ++k;
    }
    // modified Gram-Schmidt (twice, for orthogonality to working precision)
    for (int pass = 0; pass < 2; ++pass)
        for (int i = 0; i < 4; ++i) {
            for (int p = 0; p < i; ++p) {
                double dot = 0.0;
                for (int j = 0; j < 9; ++j) dot += N[i][j] * N[p][j];
                for (int j = 0; j < 9; ++j) N[i][j] -= dot * N[p][j];
            }
            double nn = 0.0;
            for (int j = 0; j < 9; ++j) nn += N[i][j] * N[i][j];
            const double inv = 1.0 / sqrt(nn);
            for (int j = 0; j < 9; ++j) N[i][j] *= inv;
        }
}

// One scaled row of the constraint matrix -> storage
template <class Mat>
GEOM_HD inline void store_row(Mat& M, int row, const double* acc) {
    double mx = 0.0;
    for (int j = 0; j < 20; ++j) mx = fmax(mx, fabs(acc[j]));
    const double inv = (mx > 0.0) ? 1.0 / mx : 1.0;     // row equilibration: pure scaling
    for (int j = 0; j < 20; ++j) M(row, j) = acc[j] * inv;
}

template <class Mat>
GEOM_HD inline bool five_point_setup(const double x1[5][2], const double x2[5][2], Mat& M, FivePointSetup& S) {
    double (*N)[9] = S.N;
    {
        double A[5][9];
        for (int i = 0; i < 5; ++i) {
            const double a = x1[i][0], b = x1[i][1], c = x2[i][0], d = x2[i][1];
            A[i][0] = c * a; A[i][1] = c * b; A[i][2] = c;
            A[i][3] = d * a; A[i][4] = d * b; A[i][5] = d;
            A[i][6] = a;     A[i][7] = b;     A[i][8] = 1.0;
        }
        null_space_5x9(A, N);
    }
    // e(k) = linear polynomial (x, y, z, 1) of entry k = r*3+c = column k of N (stride 9)
    const double* e0 = &N[0][0];
#define E_(k) (e0 + (k))
    // row 0: det E
    {
        double acc[20];
        for (int j = 0; j < 20; ++j) acc[j] = 0.0;
        for (int p = 0; p < 3; ++p) {
            const int a = p, b = (p + 1) % 3, c = (p + 2) % 3;
            double q[10];
            for (int i = 0; i < 10; ++i) q[i] = 0.0;
            lin_mul_acc(E_(3 + b), E_(6 + c), 1.0, q);       // e1b * e2c
            lin_mul_acc(E_(3 + c), E_(6 + b), -1.0, q);      // - e1c * e2b
            quad_mul_acc(q, E_(a), acc);
        }
        store_row(M, 0, acc);
    }
    // rows 1..9: (E E^T - 1/2 tr(E E^T) I) E, one row of Lambda at a time
    {
        double tr[10];
        for (int i = 0; i < 10; ++i) tr[i] = 0.0;
        for (int k = 0; k < 9; ++k) lin_mul_acc(E_(k), E_(k), 0.5, tr);
        for (int r = 0; r < 3; ++r) {
            double L[3][10];
            for (int k = 0; k < 3; ++k) {
                for (int i = 0; i < 10; ++i) L[k][i] = (k == r) ? -tr[i] : 0.0;
                for (int m = 0; m < 3; ++m) lin_mul_acc(E_(r * 3 + m), E_(k * 3 + m), 1.0, L[k]);
            }
            for (int c = 0; c < 3; ++c) {
                double acc[20];
                for (int j = 0; j < 20; ++j) acc[j] = 0.0;
                for (int k = 0; k < 3; ++k) quad_mul_acc(L[k], E_(k * 3 + c), acc);
                store_row(M, 1 + r * 3 + c, acc);
            }
        }
    }
#undef E_
    // Gaussian elimination with partial pivoting, reduced only where it is read later: the right
    // block (columns 10..19) completely, the left block only to the right of the pivot.
    for (int col = 0; col < 10; ++col) {
        int pr = col;
        double best = fabs(M(col, col));
        for (int i = col + 1; i < 10; ++i) {
            const double v = fabs(M(i, col));
            if (v > best) { best = v; pr = i; }
        }
        if (best < 1e-300) return false;
        double prow[20];
        const double inv = 1.0 / M(pr, col);
        for (int j = 0; j < 20; ++j) prow[j] = M(pr, j) * inv;
        if (pr != col)
            for (int j = 0; j < 20; ++j) M(pr, j) = M(col, j);
        for (int j = 0; j < 20; ++j) M(col, j) = prow[j];
        for (int i = 0; i < 10; ++i) {
            if (i == col) continue;
            const double f = M(i, col);
            for (int j = col + 1; j < 10; ++j) M(i, j) -= f * prow[j];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int j = 10; j < 20; ++j) M(i, j) -= f * prow[j];
        }
    }
    // B(z): rows k = e - z f, l = g - z h, m = i - z j ; entries ascending in z
    for (int r = 0; r < 3; ++r) {
        double be[10], bf[10];
        for (int j = 0; j < 10; ++j) { be[j] = M(4 + 2 * r, 10 + j); bf[j] = M(5 + 2 * r, 10 + j); }
        S.Bx[r][3] = -bf[0]; S.Bx[r][2] = be[0] - bf[1]; S.Bx[r][1] = be[1] - bf[2]; S.Bx[r][0] = be[2];
        S.By[r][3] = -bf[3]; S.By[r][2] = be[3] - bf[4]; S.By[r][1] = be[4] - bf[5]; S.By[r][0] = be[5];
        S.B1[r][4] = -bf[6]; S.B1[r][3] = be[6] - bf[7]; S.B1[r][2] = be[7] - bf[8]; S.B1[r][1] = be[8] - bf[9]; S.B1[r][0] = be[9];
    }
    // det B(z), expansion along the third column
    double* c = S.c;
    for (int i = 0; i < 11; ++i) c[i] = 0.0;
    {
        const int rows[3][2] = {{1, 2}, {0, 2}, {0, 1}};
        for (int r = 0; r < 3; ++r) {
            const int p = rows[r][0], q = rows[r][1];
            double mnr[7];
            for (int i = 0; i < 7; ++i) mnr[i] = 0.0;
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) mnr[i + j] += S.Bx[p][i] * S.By[q][j] - S.By[p][i] * S.Bx[q][j];
            const double sgn = (r == 1) ? -1.0 : 1.0;
            for (int i = 0; i < 7; ++i)
                for (int j = 0; j < 5; ++j) c[i + j] += sgn * mnr[i] * S.B1[r][j];
        }
    }
    double cmax = 0.0;
    for (int i = 0; i < 11; ++i) cmax = fmax(cmax, fabs(c[i]));
    if (!(cmax > 0.0) || !(cmax < 1e300)) return false;
    const double cinv = 1.0 / cmax;
    for (int i = 0; i < 11; ++i) c[i] *= cinv;
    return true;
}

// Brackets of the real roots: first those of p in (-1,1] ascending, then those of the reversed
// polynomial (|z| > 1) ascending in w = 1/z.  At most 10 are returned.
// allow_fallback = false: returns -1 instead of taking the slow exact path (GPU: the hypothesis is
// re-queued to a dense fallback launch so that one slow lane does not stall its whole warp).
GEOM_HD inline int five_point_isolate(const double* c, RootBracket* br, bool allow_fallback = true) {
    int n = 0;
    for (int pass = 0; pass < 2 && n < 10; ++pass) {
        double cc[11];
        for (int i = 0; i < 11; ++i) cc[i] = pass ? c[10 - i] : c[i];
        double lo[10], hi[10];
        const int k = real_root_brackets<10>(cc, 10, lo, hi, allow_fallback);
        if (k < 0) return -1;
        for (int i = 0; i < k && n < 10; ++i) {
            if (pass) {      // keep 0 < |w| < 1 only (w = 0 is z = infinity; |w| = 1 belongs to pass 0)
                if (lo[i] == hi[i] && !(fabs(lo[i]) > 1e-12 && fabs(lo[i]) < 1.0)) continue;
                if (hi[i] <= -1.0 || lo[i] >= 1.0) continue;
            }
            br[n].a = lo[i]; br[n].b = hi[i]; br[n].reversed = pass;
            ++n;
        }
    }
    return n;
}

// convenience wrapper: all three phases in one call (host tests)
GEOM_HD inline int five_point(const double x1[5][2], const double x2[5][2], double E[][9]) {
    LocalMat M;
    FivePointSetup S;
    if (!five_point_setup(x1, x2, M, S)) return 0;
    RootBracket br[10];
    const int nb = five_point_isolate(S.c, br);
    int ns = 0;
    for (int i = 0; i < nb; ++i)
        if (five_point_finish(S, br[i], E[ns])) ++ns;
    return ns;
}

}  // namespace geom
}  // namespace b200vo
