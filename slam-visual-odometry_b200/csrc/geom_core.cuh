// Two-view geometry primitives in double precision, written once for host and device.
//
// What they replace (reference call sites; the arithmetic itself lives in OpenCV):
//   cv::findEssentialMat(p1,p2,K,RANSAC,0.999,2.0,mask)   src/visual_odometry.cpp:213-216
//   cv::recoverPose(E,p1,p2,K,R,t,mask)                   src/visual_odometry.cpp:268
// Algorithms:
//   * Nister five-point: 4-D null space of the 5x9 epipolar system, the ten cubic constraints
//     det E = 0 and (E E^T - 1/2 tr(E E^T) I) E = 0 as a 10x20 matrix, Gauss-Jordan, the 3x3
//     polynomial matrix B(z), det B(z) of degree 10, its real roots, back-substitution.
//     Real roots are isolated on [-1,1] for p(z) and for the reversed polynomial (|z| > 1) by
//     recursing through the derivative chain, each bracket closed with safeguarded Newton.
//   * Sampson error exactly as OpenCV evaluates it (double, cast to float, <= float threshold).
//   * recoverPose: SVD of E by one-sided Jacobi, the four (R, t) candidates, DLT triangulation of
//     every correspondence (null vector of a 4x4 by one-sided Jacobi), cheirality + 50-baseline
//     distance test.
// The host build of this header exists ONLY so tests can check the solver against cv2 on the
// CPU box (tests/hostsim); the product calls it from CUDA kernels exclusively.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define GEOM_HD __host__ __device__
#else
#define GEOM_HD
#endif

namespace b200vo {
namespace geom {

// ----------------------------------------------------------------------------- sampler
GEOM_HD inline uint64_t splitmix_next(uint64_t& state) {
    state += 0x9E3779B97F4A7C15ull;
    uint64_t z = state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

GEOM_HD inline uint64_t pair_stream(uint64_t seed, int64_t pair_id) {
    uint64_t s = seed ^ ((uint64_t)(pair_id + 1) * 0xD1342543DE82EF95ull);
    return splitmix_next(s);
}

// 5 distinct indices in [0,m): a pure function of (seed, pair id, hypothesis index).
GEOM_HD inline void draw_sample(uint64_t base, int h, int m, int idx[5]) {
    uint64_t state = base + (uint64_t)h * 0xA24BAED4963EE407ull;
    int k = 0;
    while (k < 5) {
        const uint64_t z = splitmix_next(state);
        const int cand = (int)(((z >> 32) * (uint64_t)m) >> 32);
        bool dup = false;
        for (int j = 0; j < k; ++j) dup |= (idx[j] == cand);
        if (!dup) idx[k++] = cand;
    }
}

// ----------------------------------------------------------------------------- polynomial helpers
GEOM_HD inline void horner2(const double* c, int deg, double x, double& f, double& df) {
    f = c[deg];
    df = 0.0;
    for (int i = deg - 1; i >= 0; --i) {
        df = df * x + f;
        f = f * x + c[i];
    }
}

GEOM_HD inline double horner(const double* c, int deg, double x) {
    double f = c[deg];
    for (int i = deg - 1; i >= 0; --i) f = f * x + c[i];
    return f;
}

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
                if (ncur < MAXDEG) cur[ncur++] = bracket_root(d, dk, lo, hi, flo);
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

// ----------------------------------------------------------------------------- five-point
// lin x lin -> quad index, quad x lin -> cubic column (Nister's column order:
// x3 y3 x2y xy2 x2z x2 y2z y2 xyz xy | xz2 xz x yz2 yz y z3 z2 z 1)
GEOM_HD inline int qidx(int i, int j) {
    const int t[4][4] = {{0, 1, 2, 6}, {1, 3, 4, 7}, {2, 4, 5, 8}, {6, 7, 8, 9}};
    return t[i][j];
}
GEOM_HD inline int cidx(int q, int l) {
    const int t[10][4] = {{0, 2, 4, 5},   {2, 3, 8, 9},    {4, 8, 10, 11},  {3, 1, 6, 7},    {8, 6, 13, 14},
                          {10, 13, 16, 17}, {5, 9, 11, 12}, {9, 7, 14, 15},  {11, 14, 17, 18}, {12, 15, 18, 19}};
    return t[q][l];
}

// out(10) += s * a(4) * b(4)
GEOM_HD inline void lin_mul_acc(const double* a, const double* b, double s, double* out) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out[qidx(i, j)] += s * a[i] * b[j];
}
// out(20) += s * q(10) * l(4)
GEOM_HD inline void quad_mul_acc(const double* q, const double* l, double s, double* out) {
    for (int i = 0; i < 10; ++i)
        for (int j = 0; j < 4; ++j) out[cidx(i, j)] += s * q[i] * l[j];
}

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

// Gauss-Newton refinement of one root (x, y, z) on the ten ORIGINAL cubic constraints: removes the
// digits lost in the elimination / degree-10 expansion.  Accepted only if the residual shrinks.
GEOM_HD inline double cubic_residuals(const double M0[10][20], double x, double y, double z, double* r, double (*J)[3]) {
    const double x2 = x * x, y2 = y * y, z2 = z * z;
    const double m[20] = {x2 * x, y2 * y, x2 * y, x * y2, x2 * z, x2, y2 * z, y2, x * y * z, x * y,
                          x * z2, x * z, x, y * z2, y * z, y, z2 * z, z2, z, 1.0};
    const double dx[20] = {3 * x2, 0, 2 * x * y, y2, 2 * x * z, 2 * x, 0, 0, y * z, y, z2, z, 1, 0, 0, 0, 0, 0, 0, 0};
    const double dy[20] = {0, 3 * y2, x2, 2 * x * y, 0, 0, 2 * y * z, 2 * y, x * z, x, 0, 0, 0, z2, z, 1, 0, 0, 0, 0};
    const double dz[20] = {0, 0, 0, 0, x2, 0, y2, 0, x * y, 0, 2 * x * z, x, 0, 2 * y * z, y, 0, 3 * z2, 2 * z, 1, 0};
    double ss = 0.0;
    for (int i = 0; i < 10; ++i) {
        double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
        for (int j = 0; j < 20; ++j) {
            const double w = M0[i][j];
            a += w * m[j];
            if (J) { b += w * dx[j]; c += w * dy[j]; d += w * dz[j]; }
        }
        r[i] = a;
        ss += a * a;
        if (J) { J[i][0] = b; J[i][1] = c; J[i][2] = d; }
    }
    return ss;
}

GEOM_HD inline void polish_xyz(const double M0[10][20], double& x, double& y, double& z) {
    if (!(fabs(x) < 1e3 && fabs(y) < 1e3 && fabs(z) < 1e3)) return;
    double r[10], J[10][3];
    double ss = cubic_residuals(M0, x, y, z, r, J);
    for (int it = 0; it < 3; ++it) {
        // normal equations (3x3, symmetric) solved by Cramer's rule
        double A[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
        for (int i = 0; i < 10; ++i) {
            A[0] += J[i][0] * J[i][0]; A[1] += J[i][0] * J[i][1]; A[2] += J[i][0] * J[i][2];
            A[3] += J[i][1] * J[i][1]; A[4] += J[i][1] * J[i][2]; A[5] += J[i][2] * J[i][2];
            g[0] += J[i][0] * r[i]; g[1] += J[i][1] * r[i]; g[2] += J[i][2] * r[i];
        }
        const double c00 = A[3] * A[5] - A[4] * A[4], c01 = A[2] * A[4] - A[1] * A[5], c02 = A[1] * A[4] - A[2] * A[3];
        const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
        if (!(fabs(det) > 0.0)) return;
        const double c11 = A[0] * A[5] - A[2] * A[2], c12 = A[1] * A[2] - A[0] * A[4], c22 = A[0] * A[3] - A[1] * A[1];
        const double inv = 1.0 / det;
        const double sx = -(c00 * g[0] + c01 * g[1] + c02 * g[2]) * inv;
        const double sy = -(c01 * g[0] + c11 * g[1] + c12 * g[2]) * inv;
        const double sz = -(c02 * g[0] + c12 * g[1] + c22 * g[2]) * inv;
        const double nx = x + sx, ny = y + sy, nz = z + sz;
        const double ns = cubic_residuals(M0, nx, ny, nz, r, J);
        if (!(ns < ss)) return;           // no improvement (or NaN): keep the current root
        x = nx; y = ny; z = nz;
        if (ns < 1e-30 || ss - ns < 1e-3 * ss) return;   // converged to working precision
        ss = ns;
    }
}

// x1, x2: 5 normalised correspondences (x2^T E x1 = 0).  E: up to 10 row-major 3x3, |E|_F = 1.
GEOM_HD inline int five_point(const double x1[5][2], const double x2[5][2], double E[][9]) {
    double A[5][9], N[4][9];
    for (int i = 0; i < 5; ++i) {
        const double a = x1[i][0], b = x1[i][1], c = x2[i][0], d = x2[i][1];
        A[i][0] = c * a; A[i][1] = c * b; A[i][2] = c;
        A[i][3] = d * a; A[i][4] = d * b; A[i][5] = d;
        A[i][6] = a;     A[i][7] = b;     A[i][8] = 1.0;
    }
    null_space_5x9(A, N);
    // e[r*3+c] = linear polynomial (x, y, z, 1) of entry (r,c)
    double e[9][4];
    for (int k = 0; k < 9; ++k)
        for (int v = 0; v < 4; ++v) e[k][v] = N[v][k];

    double M[10][20];
    for (int i = 0; i < 10; ++i)
        for (int j = 0; j < 20; ++j) M[i][j] = 0.0;
    // det E
    {
        const int perm[3][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}};
        for (int p = 0; p < 3; ++p) {
            const int a = perm[p][0], b = perm[p][1], c = perm[p][2];
            double q[10];
            for (int i = 0; i < 10; ++i) q[i] = 0.0;
            lin_mul_acc(e[3 + b], e[6 + c], 1.0, q);       // e1b * e2c
            lin_mul_acc(e[3 + c], e[6 + b], -1.0, q);      // - e1c * e2b
            quad_mul_acc(q, e[a], 1.0, M[0]);
        }
    }
    // Lambda = E E^T - 1/2 tr(E E^T) I ;  rows 1..9 = Lambda E
    {
        double L[3][3][10];
        for (int r = 0; r < 3; ++r)
            for (int c = r; c < 3; ++c) {
                for (int i = 0; i < 10; ++i) L[r][c][i] = 0.0;
                for (int k = 0; k < 3; ++k) lin_mul_acc(e[r * 3 + k], e[c * 3 + k], 1.0, L[r][c]);
            }
        double tr[10];
        for (int i = 0; i < 10; ++i) tr[i] = 0.5 * (L[0][0][i] + L[1][1][i] + L[2][2][i]);
        for (int r = 0; r < 3; ++r)
            for (int i = 0; i < 10; ++i) L[r][r][i] -= tr[i];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < r; ++c)
                for (int i = 0; i < 10; ++i) L[r][c][i] = L[c][r][i];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                for (int k = 0; k < 3; ++k) quad_mul_acc(L[r][k], e[k * 3 + c], 1.0, M[1 + r * 3 + c]);
    }
    double M0[10][20];
    for (int i = 0; i < 10; ++i) {
        // equilibrate rows (pure scaling: the solution set is unchanged)
        double mx = 0.0;
        for (int j = 0; j < 20; ++j) mx = fmax(mx, fabs(M[i][j]));
        const double inv = (mx > 0.0) ? 1.0 / mx : 1.0;
        for (int j = 0; j < 20; ++j) { M[i][j] *= inv; M0[i][j] = M[i][j]; }
    }
    // Gauss-Jordan on the first 10 columns, partial pivoting
    for (int col = 0; col < 10; ++col) {
        int pr = col;
        double best = fabs(M[col][col]);
        for (int i = col + 1; i < 10; ++i)
            if (fabs(M[i][col]) > best) { best = fabs(M[i][col]); pr = i; }
        if (best < 1e-300) return 0;
        if (pr != col)
            for (int j = col; j < 20; ++j) { const double t = M[col][j]; M[col][j] = M[pr][j]; M[pr][j] = t; }
        const double inv = 1.0 / M[col][col];
        for (int j = col; j < 20; ++j) M[col][j] *= inv;
        for (int i = 0; i < 10; ++i) {
            if (i == col) continue;
            const double f = M[i][col];
            if (f != 0.0)
                for (int j = col; j < 20; ++j) M[i][j] -= f * M[col][j];
        }
    }
    // B(z): rows k = e - z f, l = g - z h, m = i - z j ; entries ascending in z
    double Bx[3][4], By[3][4], B1[3][5];
    for (int r = 0; r < 3; ++r) {
        const double* be = &M[4 + 2 * r][10];
        const double* bf = &M[5 + 2 * r][10];
        Bx[r][3] = -bf[0]; Bx[r][2] = be[0] - bf[1]; Bx[r][1] = be[1] - bf[2]; Bx[r][0] = be[2];
        By[r][3] = -bf[3]; By[r][2] = be[3] - bf[4]; By[r][1] = be[4] - bf[5]; By[r][0] = be[5];
        B1[r][4] = -bf[6]; B1[r][3] = be[6] - bf[7]; B1[r][2] = be[7] - bf[8]; B1[r][1] = be[8] - bf[9]; B1[r][0] = be[9];
    }
    // det B(z), expansion along the third column
    double c[11];
    for (int i = 0; i < 11; ++i) c[i] = 0.0;
    {
        const int rows[3][2] = {{1, 2}, {0, 2}, {0, 1}};
        for (int r = 0; r < 3; ++r) {
            const int p = rows[r][0], q = rows[r][1];
            double mnr[7];
            for (int i = 0; i < 7; ++i) mnr[i] = 0.0;
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) mnr[i + j] += Bx[p][i] * By[q][j] - By[p][i] * Bx[q][j];
            const double sgn = (r == 1) ? -1.0 : 1.0;
            for (int i = 0; i < 7; ++i)
                for (int j = 0; j < 5; ++j) c[i + j] += sgn * mnr[i] * B1[r][j];
        }
    }
    double cmax = 0.0;
    for (int i = 0; i < 11; ++i) cmax = fmax(cmax, fabs(c[i]));
    if (!(cmax > 0.0) || !(cmax < 1e300)) return 0;
    double crev[11];
    for (int i = 0; i < 11; ++i) { c[i] /= cmax; }
    for (int i = 0; i < 11; ++i) crev[i] = c[10 - i];

    double zr[20];
    int nz = real_roots_unit<10>(c, 10, zr);
    {
        double wr[10];
        const int nw = real_roots_unit<10>(crev, 10, wr);
        for (int i = 0; i < nw; ++i)
            if (fabs(wr[i]) > 1e-12 && fabs(wr[i]) < 1.0) zr[nz++] = 1.0 / wr[i];
    }
    int ns = 0;
    for (int s = 0; s < nz && ns < 10; ++s) {
        const double z = zr[s];
        double b[3][3];
        for (int r = 0; r < 3; ++r) {
            b[r][0] = horner(Bx[r], 3, z);
            b[r][1] = horner(By[r], 3, z);
            b[r][2] = horner(B1[r], 4, z);
        }
        // null vector of B(z): best-conditioned cross product of two rows
        double nbest[3] = {0, 0, 0}, nn = -1.0;
        const int pr[3][2] = {{0, 1}, {0, 2}, {1, 2}};
        for (int k = 0; k < 3; ++k) {
            const double* u = b[pr[k][0]];
            const double* v = b[pr[k][1]];
            // scale rows so the comparison is not dominated by one large row
            const double su = fabs(u[0]) + fabs(u[1]) + fabs(u[2]), sv = fabs(v[0]) + fabs(v[1]) + fabs(v[2]);
            if (!(su > 0.0) || !(sv > 0.0)) continue;
            const double cx = u[1] * v[2] - u[2] * v[1], cy = u[2] * v[0] - u[0] * v[2], cz = u[0] * v[1] - u[1] * v[0];
            const double q = (cx * cx + cy * cy + cz * cz) / (su * su * sv * sv);
            if (q > nn) { nn = q; nbest[0] = cx; nbest[1] = cy; nbest[2] = cz; }
        }
        const double nrm = sqrt(nbest[0] * nbest[0] + nbest[1] * nbest[1] + nbest[2] * nbest[2]);
        if (!(nrm > 0.0) || fabs(nbest[2]) < 1e-10 * nrm) continue;
        double x = nbest[0] / nbest[2], y = nbest[1] / nbest[2], zz = z;
        polish_xyz(M0, x, y, zz);
        double fn = 0.0;
        for (int k = 0; k < 9; ++k) {
            const double v = x * N[0][k] + y * N[1][k] + zz * N[2][k] + N[3][k];
            E[ns][k] = v;
            fn += v * v;
        }
        if (!(fn > 0.0) || !(fn < 1e300)) continue;
        const double inv = 1.0 / sqrt(fn);
        for (int k = 0; k < 9; ++k) E[ns][k] *= inv;
        ++ns;
    }
    return ns;
}

// ----------------------------------------------------------------------------- Sampson
// OpenCV EMEstimatorCallback::computeError: err = (float)(num^2/den); inlier iff err <= thr2f.
GEOM_HD inline bool sampson_inlier(const double* E, double x1, double y1, double x2, double y2, float thr2f) {
    const double a0 = E[0] * x1 + E[1] * y1 + E[2];
    const double a1 = E[3] * x1 + E[4] * y1 + E[5];
    const double a2 = E[6] * x1 + E[7] * y1 + E[8];
    const double b0 = E[0] * x2 + E[3] * y2 + E[6];
    const double b1 = E[1] * x2 + E[4] * y2 + E[7];
    const double num = x2 * a0 + y2 * a1 + a2;
    const double den = a0 * a0 + a1 * a1 + b0 * b0 + b1 * b1;
    const float err = (float)(num * num / den);
    return err <= thr2f;
}

// ----------------------------------------------------------------------------- small SVD
// One-sided (Hestenes) Jacobi: on exit the columns of G are orthogonal (G = A V), V orthogonal.
template <int N>
GEOM_HD inline void jacobi_cols(double G[N][N], double V[N][N]) {
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool changed = false;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                double al = 0.0, be = 0.0, ga = 0.0;
                for (int i = 0; i < N; ++i) { al += G[i][p] * G[i][p]; be += G[i][q] * G[i][q]; ga += G[i][p] * G[i][q]; }
                if (fabs(ga) <= 1e-16 * sqrt(al * be) || ga == 0.0) continue;
                changed = true;
                const double zeta = (be - al) / (2.0 * ga);
                const double t = ((zeta >= 0.0) ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (int i = 0; i < N; ++i) {
                    const double gp = G[i][p], gq = G[i][q];
                    G[i][p] = cs * gp - sn * gq;
                    G[i][q] = sn * gp + cs * gq;
                    const double vp = V[i][p], vq = V[i][q];
                    V[i][p] = cs * vp - sn * vq;
                    V[i][q] = sn * vp + cs * vq;
                }
            }
        if (!changed) break;
    }
}

GEOM_HD inline double det3(const double* m) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

// decomposeEssentialMat: R1 = U W Vt, R2 = U Wt Vt, t = U[:,2]   (App. A.10)
GEOM_HD inline void decompose_essential(const double* E, double* R1, double* R2, double* t) {
    double G[3][3], V[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) G[i][j] = E[i * 3 + j];
    jacobi_cols<3>(G, V);
    double s[3];
    for (int j = 0; j < 3; ++j) s[j] = sqrt(G[0][j] * G[0][j] + G[1][j] * G[1][j] + G[2][j] * G[2][j]);
    int o[3] = {0, 1, 2};
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2 - a; ++b)
            if (s[o[b]] < s[o[b + 1]]) { const int tt = o[b]; o[b] = o[b + 1]; o[b + 1] = tt; }
    double U[9], Vt[9];
    for (int j = 0; j < 2; ++j) {
        const double inv = 1.0 / s[o[j]];
        for (int i = 0; i < 3; ++i) U[i * 3 + j] = G[i][o[j]] * inv;
    }
    U[2] = U[3] * U[7] - U[6] * U[4];        // u2 = u0 x u1 (sigma_3 ~ 0 for an essential matrix)
    U[5] = U[6] * U[1] - U[0] * U[7];
    U[8] = U[0] * U[4] - U[3] * U[1];
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) Vt[j * 3 + i] = V[i][o[j]];
    if (det3(U) < 0) for (int i = 0; i < 9; ++i) U[i] = -U[i];
    if (det3(Vt) < 0) for (int i = 0; i < 9; ++i) Vt[i] = -Vt[i];
    // U W : columns (-u1, u0, u2) ; U W^T : columns (u1, -u0, u2)
    double UW[9], UWt[9];
    for (int i = 0; i < 3; ++i) {
        UW[i * 3 + 0] = -U[i * 3 + 1]; UW[i * 3 + 1] = U[i * 3 + 0];   UW[i * 3 + 2] = U[i * 3 + 2];
        UWt[i * 3 + 0] = U[i * 3 + 1]; UWt[i * 3 + 1] = -U[i * 3 + 0]; UWt[i * 3 + 2] = U[i * 3 + 2];
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double a = 0.0, b = 0.0;
            for (int k = 0; k < 3; ++k) { a += UW[i * 3 + k] * Vt[k * 3 + j]; b += UWt[i * 3 + k] * Vt[k * 3 + j]; }
            R1[i * 3 + j] = a;
            R2[i * 3 + j] = b;
        }
    t[0] = U[2]; t[1] = U[5]; t[2] = U[8];
}

// DLT triangulation with P0 = [I|0], P1 = [R|t] + the recoverPose cheirality/distance test.
GEOM_HD inline bool cheirality_ok(const double* R, const double* t, double x1, double y1, double x2, double y2,
                                  double dist) {
    double G[4][4], V[4][4];
    G[0][0] = -1.0; G[0][1] = 0.0;  G[0][2] = x1; G[0][3] = 0.0;
    G[1][0] = 0.0;  G[1][1] = -1.0; G[1][2] = y1; G[1][3] = 0.0;
    for (int k = 0; k < 3; ++k) {
        G[2][k] = x2 * R[6 + k] - R[k];
        G[3][k] = y2 * R[6 + k] - R[3 + k];
    }
    G[2][3] = x2 * t[2] - t[0];
    G[3][3] = y2 * t[2] - t[1];
    jacobi_cols<4>(G, V);
    int jm = 0;
    double smin = 1e300;
    for (int j = 0; j < 4; ++j) {
        const double s = G[0][j] * G[0][j] + G[1][j] * G[1][j] + G[2][j] * G[2][j] + G[3][j] * G[3][j];
        if (s < smin) { smin = s; jm = j; }
    }
    const double Q0 = V[0][jm], Q1 = V[1][jm], Q2 = V[2][jm], Q3 = V[3][jm];
    bool ok = (Q2 * Q3 > 0.0);
    const double X = Q0 / Q3, Y = Q1 / Q3, Z = Q2 / Q3;
    ok = ok && (Z < dist);
    const double z2 = R[6] * X + R[7] * Y + R[8] * Z + t[2];
    ok = ok && (z2 > 0.0) && (z2 < dist);
    return ok;
}

}  // namespace geom
}  // namespace b200vo
