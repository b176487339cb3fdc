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
// The five-point solver itself is in fivept_group.cuh (cooperative, 10 lanes per hypothesis); this
// header holds what it shares with the per-root finish, the scorer and recoverPose.
// The host build of this header exists ONLY so tests can check the code against cv2 on the
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

// Newton inside a sign-changing bracket, bisection when unsafe; stops as soon as the step is small
// enough that one more Newton step reaches working precision (quadratic convergence).
GEOM_HD inline double bracket_root_fast(const double* c, int deg, double a, double b, double fa) {
    double x = 0.5 * (a + b);
    for (int it = 0; it < 64; ++it) {
        double f, df;
        horner2(c, deg, x, f, df);
        if (f == 0.0) return x;
        if ((f < 0.0) == (fa < 0.0)) a = x; else b = x;
        double xn = x - f / df;
        const bool newton_ok = (xn > a && xn < b);
        if (!newton_ok) xn = 0.5 * (a + b);
        const double step = fabs(xn - x);
        x = xn;
        if (newton_ok && step < 1e-6 * (fabs(x) + 1e-3)) {      // next step ~ step^2: polish once and stop
            horner2(c, deg, x, f, df);
            const double x2 = x - f / df;
            return (x2 > a && x2 < b) ? x2 : x;
        }
        if (b - a <= 4e-16 * fabs(x) + 1e-300) return x;
    }
    return x;
}

// The same with the coefficients of a degree <= 10 polynomial in REGISTERS (c[i] = 0 above the degree) and every
// loop unrolled.  The exact root path evaluates polynomials of run-time degree from shared memory: one dependent
// load + FMA per Horner step (~35 cycles) inside up to 64 Newton steps, ten levels deep -- the slowest warp of the
// streaming step.  Leading zero coefficients leave every intermediate value unchanged (0 * x + c = c exactly), so the
// results are bit-identical to horner2 / bracket_root_fast on the true degree.
GEOM_HD inline void horner2_reg10(const double* c, double x, double& f, double& df) {
    f = c[10];
    df = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 9; i >= 0; --i) {
        df = df * x + f;
        f = f * x + c[i];
    }
}
GEOM_HD inline double horner_reg10(const double* c, double x) {
    double f = c[10];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 9; i >= 0; --i) f = f * x + c[i];
    return f;
}
GEOM_HD inline double bracket_root_fast_reg10(const double* c, double a, double b, double fa) {
    double x = 0.5 * (a + b);
    for (int it = 0; it < 64; ++it) {
        double f, df;
        horner2_reg10(c, x, f, df);
        if (f == 0.0) return x;
        if ((f < 0.0) == (fa < 0.0)) a = x; else b = x;
        double xn = x - f / df;
        const bool newton_ok = (xn > a && xn < b);
        if (!newton_ok) xn = 0.5 * (a + b);
        const double step = fabs(xn - x);
        x = xn;
        if (newton_ok && step < 1e-6 * (fabs(x) + 1e-3)) {
            horner2_reg10(c, x, f, df);
            const double x2 = x - f / df;
            return (x2 > a && x2 < b) ? x2 : x;
        }
        if (b - a <= 4e-16 * fabs(x) + 1e-300) return x;
    }
    return x;
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

// out(10) += s * a(4) * b(4)      (fully unrolled: every index is a compile-time constant)
// a, b: linear polynomials stored with stride 9 (column k of the 4x9 null-space basis)
GEOM_HD inline void lin_mul_acc(const double* a, const double* b, double s, double* out) {
    double bb[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 4; ++j) bb[j] = b[9 * j];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; ++i) {
        const double sa = s * a[9 * i];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 4; ++j) out[qidx(i, j)] += sa * bb[j];
    }
}
// out(20) += q(10) * l(4)
GEOM_HD inline void quad_mul_acc(const double* q, const double* l, double* out) {
    double ll[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 4; ++j) ll[j] = l[9 * j];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 10; ++i) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 4; ++j) out[cidx(i, j)] += q[i] * ll[j];
    }
}

// Residuals of the ten cubic constraints at a numeric E, r[0] = det E, r[1..9] = (E E^T - 1/2 tr I) E.
// With D != nullptr also the directional derivative dr along dE = D (all 3x3 row-major).
GEOM_HD inline void essential_residuals(const double* E, double* r) {
    double G[9];                      // Lambda = E E^T - 1/2 tr(E E^T) I
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) G[i * 3 + j] = E[i * 3] * E[j * 3] + E[i * 3 + 1] * E[j * 3 + 1] + E[i * 3 + 2] * E[j * 3 + 2];
    const double ht = 0.5 * (G[0] + G[4] + G[8]);
    G[0] -= ht; G[4] -= ht; G[8] -= ht;
    r[0] = E[0] * (E[4] * E[8] - E[5] * E[7]) - E[1] * (E[3] * E[8] - E[5] * E[6]) + E[2] * (E[3] * E[7] - E[4] * E[6]);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r[1 + i * 3 + j] = G[i * 3] * E[j] + G[i * 3 + 1] * E[3 + j] + G[i * 3 + 2] * E[6 + j];
}

GEOM_HD inline void essential_residual_derivative(const double* E, const double* D, double* dr) {
    // d det = sum cof(E)_ij D_ij
    const double c0 = E[4] * E[8] - E[5] * E[7], c1 = E[5] * E[6] - E[3] * E[8], c2 = E[3] * E[7] - E[4] * E[6];
    const double c3 = E[2] * E[7] - E[1] * E[8], c4 = E[0] * E[8] - E[2] * E[6], c5 = E[1] * E[6] - E[0] * E[7];
    const double c6 = E[1] * E[5] - E[2] * E[4], c7 = E[2] * E[3] - E[0] * E[5], c8 = E[0] * E[4] - E[1] * E[3];
    dr[0] = c0 * D[0] + c1 * D[1] + c2 * D[2] + c3 * D[3] + c4 * D[4] + c5 * D[5] + c6 * D[6] + c7 * D[7] + c8 * D[8];
    // d[(EE^T - 1/2 tr I)E] = (D E^T + E D^T - tr(E D^T) I) E + (E E^T - 1/2 tr(EE^T) I) D
    double S[9], G[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            S[i * 3 + j] = D[i * 3] * E[j * 3] + D[i * 3 + 1] * E[j * 3 + 1] + D[i * 3 + 2] * E[j * 3 + 2] +
                           E[i * 3] * D[j * 3] + E[i * 3 + 1] * D[j * 3 + 1] + E[i * 3 + 2] * D[j * 3 + 2];
            G[i * 3 + j] = E[i * 3] * E[j * 3] + E[i * 3 + 1] * E[j * 3 + 1] + E[i * 3 + 2] * E[j * 3 + 2];
        }
    const double hs = 0.5 * (S[0] + S[4] + S[8]);      // tr(E D^T) = 1/2 tr(S)
    S[0] -= hs; S[4] -= hs; S[8] -= hs;
    const double ht = 0.5 * (G[0] + G[4] + G[8]);
    G[0] -= ht; G[4] -= ht; G[8] -= ht;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            dr[1 + i * 3 + j] = S[i * 3] * E[j] + S[i * 3 + 1] * E[3 + j] + S[i * 3 + 2] * E[6 + j] +
                                G[i * 3] * D[j] + G[i * 3 + 1] * D[3 + j] + G[i * 3 + 2] * D[6 + j];
}

// Gauss-Newton refinement of one root on the ten ORIGINAL cubic constraints, evaluated at
// E = u0 N0 + u1 N1 + u2 N2 + u3 N3.  The root arrives as u = (x, y, z, 1); the component of
// largest magnitude is fixed to 1 and the other three are refined, so the iteration is well scaled
// for roots near infinity too (|z| >> 1 comes from the reversed polynomial).  This removes the
// digits lost in the elimination / degree-10 expansion.  A step is accepted only if the residual
// shrinks; at most 4 steps, stopping at |r| ~ 1e-13.
GEOM_HD inline void polish_u(const double N[4][9], double* u) {
    int f = 3;
    for (int i = 0; i < 3; ++i) if (fabs(u[i]) > fabs(u[f])) f = i;
    if (!(fabs(u[f]) > 0.0) || !(fabs(u[f]) < 1e300)) return;
    // The fixed component is moved to slot 3 (rows of N are addressed through offsets), so every
    // array index below is a compile-time constant -- no run-time register indexing on the GPU.
    double w0 = u[0], w1 = u[1], w2 = u[2], w3 = u[3];
    int o0 = 0, o1 = 9, o2 = 18, o3 = 27;
    if (f == 0) { const double t = w0; w0 = w3; w3 = t; o0 = 27; o3 = 0; }
    else if (f == 1) { const double t = w1; w1 = w3; w3 = t; o1 = 27; o3 = 9; }
    else if (f == 2) { const double t = w2; w2 = w3; w3 = t; o2 = 27; o3 = 18; }
    {
        const double inv = 1.0 / w3;
        w0 *= inv; w1 *= inv; w2 *= inv; w3 = 1.0;
    }
    const double* Nf = &N[0][0];
    double E[9], r[10];
    for (int k = 0; k < 9; ++k) E[k] = w0 * Nf[o0 + k] + w1 * Nf[o1 + k] + w2 * Nf[o2 + k] + Nf[o3 + k];
    essential_residuals(E, r);
    double ss = 0.0;
    for (int i = 0; i < 10; ++i) ss += r[i] * r[i];
    for (int it = 0; it < 4; ++it) {
        if (ss < 1e-26) break;            // |r| ~ 1e-13: already at working precision
        double J0[10], J1[10], J2[10];
        essential_residual_derivative(E, Nf + o0, J0);
        essential_residual_derivative(E, Nf + o1, J1);
        essential_residual_derivative(E, Nf + o2, J2);
        // normal equations (3x3, symmetric) solved by Cramer's rule
        double A0 = 0, A1 = 0, A2 = 0, A3 = 0, A4 = 0, A5 = 0, g0 = 0, g1 = 0, g2 = 0;
        for (int i = 0; i < 10; ++i) {
            A0 += J0[i] * J0[i]; A1 += J0[i] * J1[i]; A2 += J0[i] * J2[i];
            A3 += J1[i] * J1[i]; A4 += J1[i] * J2[i]; A5 += J2[i] * J2[i];
            g0 += J0[i] * r[i]; g1 += J1[i] * r[i]; g2 += J2[i] * r[i];
        }
        const double c00 = A3 * A5 - A4 * A4, c01 = A2 * A4 - A1 * A5, c02 = A1 * A4 - A2 * A3;
        const double det = A0 * c00 + A1 * c01 + A2 * c02;
        if (!(fabs(det) > 0.0)) break;
        const double c11 = A0 * A5 - A2 * A2, c12 = A1 * A2 - A0 * A4, c22 = A0 * A3 - A1 * A1;
        const double inv = 1.0 / det;
        const double n0 = w0 - (c00 * g0 + c01 * g1 + c02 * g2) * inv;
        const double n1 = w1 - (c01 * g0 + c11 * g1 + c12 * g2) * inv;
        const double n2 = w2 - (c02 * g0 + c12 * g1 + c22 * g2) * inv;
        double En[9], rn[10];
        for (int k = 0; k < 9; ++k) En[k] = n0 * Nf[o0 + k] + n1 * Nf[o1 + k] + n2 * Nf[o2 + k] + Nf[o3 + k];
        essential_residuals(En, rn);
        double ns = 0.0;
        for (int i = 0; i < 10; ++i) ns += rn[i] * rn[i];
        if (!(ns < ss)) break;            // no improvement (or NaN): keep the current root
        w0 = n0; w1 = n1; w2 = n2;
        for (int k = 0; k < 9; ++k) E[k] = En[k];
        for (int i = 0; i < 10; ++i) r[i] = rn[i];
        if (ss - ns < 1e-3 * ss) break;   // stalled
        ss = ns;
    }
    // undo the permutation
    if (f == 0) { const double t = w0; w0 = w3; w3 = t; }
    else if (f == 1) { const double t = w1; w1 = w3; w3 = t; }
    else if (f == 2) { const double t = w2; w2 = w3; w3 = t; }
    u[0] = w0; u[1] = w1; u[2] = w2; u[3] = w3;
}

// ---- what the solver (fivept_group.cuh) hands to the per-root finish ----
struct FivePointSetup {          // 86 doubles, kept flat so it can be stored element-major
    double N[4][9];              // null-space basis: E = x N0 + y N1 + z N2 + N3
    double Bx[3][4], By[3][4], B1[3][5];   // B(z), ascending powers of z
    double c[11];                // det B(z), normalised to max |c| = 1
};
constexpr int kSetupDoubles = 86;

struct RootBracket {
    double a, b;                 // one root in (a, b]; a == b: already refined
    int reversed;                // 1: root w of the reversed polynomial, z = 1/w
};

// One bracket -> one essential matrix (|E|_F = 1).  false: no valid solution from this root.
GEOM_HD inline bool five_point_finish(const FivePointSetup& S, const RootBracket& br, double* Es) {
    double z;
    {
        double cc[11];
        for (int i = 0; i < 11; ++i) cc[i] = br.reversed ? S.c[10 - i] : S.c[i];
        double w = br.a;
        if (br.a != br.b) w = bracket_root_fast(cc, 10, br.a, br.b, horner(cc, 10, br.a));
        if (br.reversed) {
            if (!(fabs(w) > 1e-12 && fabs(w) < 1.0)) return false;
            z = 1.0 / w;
        } else {
            z = w;
        }
    }
    double b[3][3];
    for (int r = 0; r < 3; ++r) {
        b[r][0] = horner(S.Bx[r], 3, z);
        b[r][1] = horner(S.By[r], 3, z);
        b[r][2] = horner(S.B1[r], 4, z);
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
    if (!(nrm > 0.0) || fabs(nbest[2]) < 1e-10 * nrm) return false;
    double u[4] = {nbest[0] / nbest[2], nbest[1] / nbest[2], z, 1.0};
    polish_u(S.N, u);
    double fn = 0.0;
    for (int k = 0; k < 9; ++k) {
        const double v = u[0] * S.N[0][k] + u[1] * S.N[1][k] + u[2] * S.N[2][k] + u[3] * S.N[3][k];
        Es[k] = v;
        fn += v * v;
    }
    if (!(fn > 0.0) || !(fn < 1e300)) return false;
    const double inv = 1.0 / sqrt(fn);
    for (int k = 0; k < 9; ++k) Es[k] *= inv;
    return true;
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

// The same predicate without the division on the common path.  (float)q <= thr2f  <=>  q <= T',
// T' = the midpoint between thr2f and the next float above it (round-to-nearest boundary), so the
// test is num^2 <= T' * den; inside a 1e-14 relative guard band around equality the exact
// expression (with the division and the float cast) decides, which keeps the mask identical.
struct SampsonThreshold {
    float thr2f;
    double tmid;      // T'
};
GEOM_HD inline SampsonThreshold make_sampson_threshold(float thr2f) {
    SampsonThreshold t;
    t.thr2f = thr2f;
    t.tmid = 0.5 * ((double)thr2f + (double)nextafterf(thr2f, 3.0e38f));
    return t;
}
GEOM_HD inline bool sampson_inlier_fast(const double* E, double x1, double y1, double x2, double y2, const SampsonThreshold& t) {
    const double a0 = E[0] * x1 + E[1] * y1 + E[2];
    const double a1 = E[3] * x1 + E[4] * y1 + E[5];
    const double a2 = E[6] * x1 + E[7] * y1 + E[8];
    const double b0 = E[0] * x2 + E[3] * y2 + E[6];
    const double b1 = E[1] * x2 + E[4] * y2 + E[7];
    const double num = x2 * a0 + y2 * a1 + a2;
    const double den = a0 * a0 + a1 * a1 + b0 * b0 + b1 * b1;
    const double lhs = num * num, rhs = t.tmid * den;
    const double band = 1e-14 * rhs;
    if (lhs < rhs - band) return true;
    if (lhs > rhs + band) return false;
    return (float)(lhs / den) <= t.thr2f;       // also NaN / den == 0 land here
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

// Smallest right singular vector of a 4x4 matrix G (the DLT solution cv::triangulatePoints takes from an SVD)
// without an iterative SVD: it is the eigenvector of A = G^T G for the smallest eigenvalue, and
// adj(A) = sum_i (prod_{j != i} lambda_j) v_i v_i^T is dominated by that eigenvector by the factor lambda_2 / lambda_1
// = (sigma_3 / sigma_4)^2 (parallax over noise, squared: >= 1e3 for a triangulable point).  So: the column of adj(A)
// with the largest diagonal entry, multiplied by adj(A) three more times -- ~250 multiply-adds, no square root, no
// division.  Convergence is CHECKED (the last two iterates must be parallel to 1e-12); a point that fails the check
// (sigma_3 ~ sigma_4: no parallax at all) takes the one-sided Jacobi SVD, so the result is the least-squares vector
// in every case.  (Per-thread Jacobi was 47 us of the streaming step's critical path and 0.7 ms per 999 pairs.)
GEOM_HD inline bool null_vector4_adjugate(const double G[4][4], double* Q) {
    double m[4][4];
    for (int i = 0; i < 4; ++i)
        for (int j = i; j < 4; ++j) {
            const double v = G[0][i] * G[0][j] + G[1][i] * G[1][j] + G[2][i] * G[2][j] + G[3][i] * G[3][j];
            m[i][j] = v; m[j][i] = v;
        }
    const double s0 = m[0][0] * m[1][1] - m[1][0] * m[0][1], s1 = m[0][0] * m[1][2] - m[1][0] * m[0][2];
    const double s2 = m[0][0] * m[1][3] - m[1][0] * m[0][3], s3 = m[0][1] * m[1][2] - m[1][1] * m[0][2];
    const double s4 = m[0][1] * m[1][3] - m[1][1] * m[0][3], s5 = m[0][2] * m[1][3] - m[1][2] * m[0][3];
    const double c5 = m[2][2] * m[3][3] - m[3][2] * m[2][3], c4 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    const double c3 = m[2][1] * m[3][2] - m[3][1] * m[2][2], c2 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    const double c1 = m[2][0] * m[3][2] - m[3][0] * m[2][2], c0 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    double b[4][4];                      // adj(A), symmetric
    b[0][0] = m[1][1] * c5 - m[1][2] * c4 + m[1][3] * c3;
    b[0][1] = -m[0][1] * c5 + m[0][2] * c4 - m[0][3] * c3;
    b[0][2] = m[3][1] * s5 - m[3][2] * s4 + m[3][3] * s3;
    b[0][3] = -m[2][1] * s5 + m[2][2] * s4 - m[2][3] * s3;
    b[1][1] = m[0][0] * c5 - m[0][2] * c2 + m[0][3] * c1;
    b[1][2] = -m[3][0] * s5 + m[3][2] * s2 - m[3][3] * s1;
    b[1][3] = m[2][0] * s5 - m[2][2] * s2 + m[2][3] * s1;
    b[2][2] = m[3][0] * s4 - m[3][1] * s2 + m[3][3] * s0;
    b[2][3] = -m[2][0] * s4 + m[2][1] * s2 - m[2][3] * s0;
    b[3][3] = m[2][0] * s3 - m[2][1] * s1 + m[2][2] * s0;
    b[1][0] = b[0][1]; b[2][0] = b[0][2]; b[3][0] = b[0][3]; b[2][1] = b[1][2]; b[3][1] = b[1][3]; b[3][2] = b[2][3];
    // start: column with the largest diagonal entry (diag(adj) = mu_1 v_k^2 + ...), selected without run-time indexing
    double x0 = b[0][0], x1 = b[1][0], x2 = b[2][0], x3 = b[3][0], dmax = b[0][0];
    if (b[1][1] > dmax) { dmax = b[1][1]; x0 = b[0][1]; x1 = b[1][1]; x2 = b[2][1]; x3 = b[3][1]; }
    if (b[2][2] > dmax) { dmax = b[2][2]; x0 = b[0][2]; x1 = b[1][2]; x2 = b[2][2]; x3 = b[3][2]; }
    if (b[3][3] > dmax) { dmax = b[3][3]; x0 = b[0][3]; x1 = b[1][3]; x2 = b[2][3]; x3 = b[3][3]; }
    if (!(dmax > 0.0)) return false;
    // scale adj so that repeated products stay in range (an exact power of two would be nicer; any scale works)
    const double sc = 1.0 / dmax;
    double y0 = 0, y1 = 0, y2 = 0, y3 = 0;
    for (int it = 0; it < 3; ++it) {
        y0 = (b[0][0] * x0 + b[0][1] * x1 + b[0][2] * x2 + b[0][3] * x3) * sc;
        y1 = (b[1][0] * x0 + b[1][1] * x1 + b[1][2] * x2 + b[1][3] * x3) * sc;
        y2 = (b[2][0] * x0 + b[2][1] * x1 + b[2][2] * x2 + b[2][3] * x3) * sc;
        y3 = (b[3][0] * x0 + b[3][1] * x1 + b[3][2] * x2 + b[3][3] * x3) * sc;
        if (it < 2) { x0 = y0; x1 = y1; x2 = y2; x3 = y3; }
    }
    // x = second-to-last iterate, y = last: parallel to 1e-12 of |x| |y| ?
    const double nx = fmax(fmax(fabs(x0), fabs(x1)), fmax(fabs(x2), fabs(x3)));
    const double ny = fmax(fmax(fabs(y0), fabs(y1)), fmax(fabs(y2), fabs(y3)));
    const double tol = 1e-12 * nx * ny;
    const bool par = fabs(x0 * y1 - x1 * y0) <= tol && fabs(x0 * y2 - x2 * y0) <= tol && fabs(x0 * y3 - x3 * y0) <= tol &&
                     fabs(x1 * y2 - x2 * y1) <= tol && fabs(x1 * y3 - x3 * y1) <= tol && fabs(x2 * y3 - x3 * y2) <= tol;
    if (!par || !(ny > 0.0) || !(ny < 1e300)) return false;
    Q[0] = y0; Q[1] = y1; Q[2] = y2; Q[3] = y3;
    return true;
}

// The rare path of cheirality_ok (no parallax: the adjugate iteration did not converge): one-sided Jacobi SVD.
GEOM_HD inline void null_vector4_jacobi(double G[4][4], double* Q) {
    double V[4][4];
    jacobi_cols<4>(G, V);
    int jm = 0;
    double smin = 1e300;
    for (int j = 0; j < 4; ++j) {
        const double s = G[0][j] * G[0][j] + G[1][j] * G[1][j] + G[2][j] * G[2][j] + G[3][j] * G[3][j];
        if (s < smin) { smin = s; jm = j; }
    }
    Q[0] = V[0][jm]; Q[1] = V[1][jm]; Q[2] = V[2][jm]; Q[3] = V[3][jm];
}

// DLT triangulation with P0 = [I|0], P1 = [R|t] + the recoverPose cheirality/distance test.
GEOM_HD inline bool cheirality_ok(const double* R, const double* t, double x1, double y1, double x2, double y2,
                                  double dist) {
    double G[4][4];
    G[0][0] = -1.0; G[0][1] = 0.0;  G[0][2] = x1; G[0][3] = 0.0;
    G[1][0] = 0.0;  G[1][1] = -1.0; G[1][2] = y1; G[1][3] = 0.0;
    for (int k = 0; k < 3; ++k) {
        G[2][k] = x2 * R[6 + k] - R[k];
        G[3][k] = y2 * R[6 + k] - R[3 + k];
    }
    G[2][3] = x2 * t[2] - t[0];
    G[3][3] = y2 * t[2] - t[1];
    double Q[4];
    if (!null_vector4_adjugate(G, Q)) null_vector4_jacobi(G, Q);
    const double Q0 = Q[0], Q1 = Q[1], Q2 = Q[2], Q3 = Q[3];
    bool ok = (Q2 * Q3 > 0.0);
    const double X = Q0 / Q3, Y = Q1 / Q3, Z = Q2 / Q3;
    ok = ok && (Z < dist);
    const double z2 = R[6] * X + R[7] * Y + R[8] * Z + t[2];
    ok = ok && (z2 > 0.0) && (z2 < dist);
    return ok;
}

// test hook: 1 when the adjugate iteration converged for this point (0: the Jacobi path was taken)
GEOM_HD inline int cheirality_fast_path(const double* R, const double* t, double x1, double y1, double x2, double y2) {
    double G[4][4], Q[4];
    G[0][0] = -1.0; G[0][1] = 0.0;  G[0][2] = x1; G[0][3] = 0.0;
    G[1][0] = 0.0;  G[1][1] = -1.0; G[1][2] = y1; G[1][3] = 0.0;
    for (int k = 0; k < 3; ++k) { G[2][k] = x2 * R[6 + k] - R[k]; G[3][k] = y2 * R[6 + k] - R[3 + k]; }
    G[2][3] = x2 * t[2] - t[0];
    G[3][3] = y2 * t[2] - t[1];
    return null_vector4_adjugate(G, Q) ? 1 : 0;
}

}  // namespace geom
}  // namespace b200vo
