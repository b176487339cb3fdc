// Five-point minimal solver, cooperative form: ONE hypothesis per group of 10 lanes.
//
// What it replaces: the minimal solver inside cv::findEssentialMat(p1, p2, K, RANSAC, ...)
// (reference src/visual_odometry.cpp:213-216); same algorithm as geom_core.cuh::five_point
// (Nister: null space, ten cubic constraints, elimination, det B(z), real roots, back-substitution,
// polish), re-laid-out for the GPU.
//
// Why: the thread-per-hypothesis version keeps a 10x20 matrix, a 5x9 matrix and a Sturm chain in
// per-thread local memory with run-time row/column indices (pivoting) -- 5.6 KB per thread, 85 % of
// its cycles stalled on local-memory traffic (profiles/ncu_r01_final_kernels.txt).  Here
//   * the 10x20 constraint matrix lives in the REGISTERS of 10 lanes (lane = row, 20 doubles each);
//     pivot rows travel through shared memory, every register index is a compile-time constant;
//   * the 5x9 epipolar system is eliminated with lane = column (9 lanes, complete pivoting), the
//     basis is orthonormalised with a Cholesky-QR (== Gram-Schmidt, same Q) from a 4x4 Gram matrix;
//   * the two Sturm chains (p and the reversed polynomial) are built by the two half-groups and
//     the roots are isolated breadth-first: every lane bisects one pending interval per round, so
//     the data-dependent depth-first recursion (8 of 32 lanes active) becomes lock-step rounds;
//   * the <= 10 root brackets of the hypothesis are finished by one lane each (five_point_finish,
//     setup data read from shared memory).
// Three groups share a warp (30 of 32 lanes) and advance in lock step (full-warp barrier per phase).
// All inter-lane exchange goes through the GroupShared block -- no shuffles -- so the SAME source
// runs on the host (tests/hostsim: the lanes of a phase are executed one after another) and is
// checked there against the cv2 golden vectors before it ever reaches a GPU.
//
// Rule for every phase (lambda passed to Exec::run): a lane may read shared data written in an
// EARLIER phase and write shared data read in a LATER phase, never both in one phase (atomics
// excepted).  Exec::run ends with the group barrier.
#pragma once
#include <cstring>
#include "geom_core.cuh"

namespace b200vo {
namespace geom {

constexpr int kGroupLanes = 10;

#if defined(__CUDACC__)
#define GEOM_HDI __host__ __device__ __forceinline__
#else
#define GEOM_HDI inline
#endif

struct PendInterval {
    double a, b;
    int pk;              // Sturm sign changes at a, b and the sign of p at a, b, packed (pack_counts)
    int pad_;
};
constexpr int kPendSlots = 6;         // pending intervals hold >= 2 roots each (or one unconfirmed root)

// Per-hypothesis block of shared memory.  Regions whose lifetimes do not overlap share storage:
// `ex` (phases A-D) with the isolation work lists (phase E), the Sturm recurrence (phase E) with the
// final bracket list (end of E, F).  2672 bytes = 112 (mod 128): the three groups of a warp touch
// disjoint banks when they read the same member.
struct GroupShared {
    FivePointSetup S;                 // N, Bx, By, B1, c -- filled by phases A-D, read by the finish phase
    union {
        double ex[120];               // exchange scratch, meaning is phase-local
        struct {
            PendInterval pend[2][3][kPendSlots];   // [polynomial][buffer][slot]
            double cand_a[2][10], cand_b[2][10];   // isolated brackets per polynomial, unordered
        } iso;
    } u;
    union {
        double chain[2][3][10];       // Sturm recurrence [polynomial][a|b|s][k], k = 1..9
        struct { double a[10], b[10]; int rev[10]; } fin;   // final ordered bracket list of the hypothesis
    } v;
    int npend[2][3];
    int ncand[2];
    int ev_v[2][5], ev_s[2][5];
    unsigned key[16];                 // pivot keys: [0..4] null space, [5..14] elimination
    int ibuf[12];
    int xflag[2][10];                 // exact path: which interval of the current / previous level holds a root
    int chain_ok[2];
    int fail;
    int nfinal;
};

GEOM_HD inline int lane_slot(int r) {
#if defined(__CUDA_ARCH__)
    (void)r;
    return 0;                         // device: lane state is per-thread, arrays have one slot
#else
    return r;                         // host emulation: one slot per lane
#endif
}
#if defined(__CUDA_ARCH__)
#define B200VO_GRP_SLOTS 1
#else
#define B200VO_GRP_SLOTS kGroupLanes
#endif

// The barrier after every phase is a FULL-warp barrier: the three groups of a warp (and the two
// spare lanes) walk through the phases in lock step, so one issued instruction serves all 30
// lanes.  (Measured: with per-group __syncwarp(mask) the hardware schedules the groups as three
// separate sub-warps -- 10.6 active lanes per instruction, three times the issue slots.)  Hence a
// group never leaves early: a dead or finished group idles through the remaining phases, and
// data-dependent loops run while ANY group of the warp still has work (Exec::any).
struct HostGroupExec {
    bool active = true;
    template <class F> void run(F&& f) { if (active) for (int r = 0; r < kGroupLanes; ++r) f(r); }
    bool any(bool c) const { return active && c; }
};
#if defined(__CUDACC__)
struct DeviceGroupExec {
    int r;
    bool active;
    template <class F> __device__ __forceinline__ void run(F&& f) { if (active) f(r); __syncwarp(); }
    __device__ __forceinline__ bool any(bool c) const { return __any_sync(0xffffffffu, active && c) != 0; }
};
#endif

GEOM_HD inline void grp_atomic_max(unsigned* p, unsigned v) {
#if defined(__CUDA_ARCH__)
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}
GEOM_HD inline int grp_atomic_inc(int* p) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, 1);
#else
    return (*p)++;
#endif
}
// order-preserving 32-bit key of |v| (float precision is plenty to choose a pivot); low 7 bits free
GEOM_HD inline unsigned mag_key(double v) {
    const float f = (float)fabs(v);
    unsigned u;
#if defined(__CUDA_ARCH__)
    u = __float_as_uint(f);
#else
    std::memcpy(&u, &f, 4);
#endif
    return u & 0xFFFFFF80u;
}
GEOM_HD inline double inv_sqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}
// 2^-e for x = f * 2^e, f in [1,2): an exact positive rescaling factor without a division
GEOM_HD inline double pow2_inv_scale(double x) {
#if defined(__CUDA_ARCH__)
    const int hi = __double2hiint(x);
    const int e = (hi >> 20) & 0x7FF;
    return __hiloint2double((2046 - e) << 20, 0);
#else
    return ldexp(1.0, -ilogb(x));
#endif
}

GEOM_HD inline int pack_counts(int ca, int cb, int sa, int sb) { return (ca & 0xFF) | ((cb & 0xFF) << 8) | ((sa + 1) << 16) | ((sb + 1) << 18); }
// high word of |x|: orders like |x| itself, so the largest magnitude of a set is an integer max
GEOM_HD inline int hi_abs(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x) & 0x7FFFFFFF;
#else
    unsigned long long u;
    std::memcpy(&u, &x, 8);
    return (int)((u >> 32) & 0x7FFFFFFFull);
#endif
}
GEOM_HD inline double pow2_from_hi(int hi) {          // 2^-e for the exponent field of a high word
    const int e = (hi >> 20) & 0x7FF;
#if defined(__CUDA_ARCH__)
    return __hiloint2double((2046 - e) << 20, 0);
#else
    return ldexp(1.0, 1023 - e);
#endif
}

// Sturm chain of a degree-10 polynomial with every index a compile-time constant (registers).
// Same recurrence as SturmChain<10>::build; remainders are rescaled by an exact power of two taken
// from the largest high word (integer max instead of a floating-point max-abs scan).
GEOM_HDI bool sturm_build10(const double* c, double* ca, double* cb, double* cs) {
    double f[11], g[10];
    double gscale = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i <= 10; ++i) f[i] = c[i];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 10; ++i) { g[i] = c[i + 1] * (double)(i + 1); gscale = fmax(gscale, fabs(g[i])); }
    bool ok = gscale > 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 1; k < 10; ++k) {
        const int m = 11 - k;                 // deg f = m, deg g = m - 1
        const double lead = g[m - 1];
        // a remainder is rescaled to max |coefficient| in [1,2), so its size is known without a scan
        ok = ok && (fabs(lead) > 1e-13 * gscale);
        const double inv = 1.0 / lead;
        const double ak = f[m] * inv;
        const double bk = (f[m - 1] - ak * ((m >= 2) ? g[m - 2] : 0.0)) * inv;
        int hmax = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i <= m - 2; ++i) {
            const double gi1 = (i >= 1) ? g[i - 1] : 0.0;
            const double r = f[i] - ak * gi1 - bk * g[i];
            f[i] = -r;
            const int hr = hi_abs(r);
            hmax = (hr > hmax) ? hr : hmax;
        }
        ok = ok && (hmax >= 0x00100000) && (hmax < 0x7E300000);      // a normal number below ~1e300
        const double sk = ok ? pow2_from_hi(hmax) : 1.0;
        ca[k] = ak; cb[k] = bk; cs[k] = sk;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i <= m - 2; ++i) { const double t = f[i] * sk; f[i] = g[i]; g[i] = t; }
        f[m - 1] = g[m - 1];
        gscale = 1.0;
    }
    return ok;
}

// sign changes of the chain at x, and the sign of p(x); p in registers, recurrence from shared memory
GEOM_HDI int sturm_eval10(const double* p, const double* ca, const double* cb, const double* cs, double x, int& sgn_p) {
    double v0 = p[10], v1 = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 9; i >= 0; --i) { v1 = v1 * x + v0; v0 = v0 * x + p[i]; }      // p(x) and p'(x)
    int cnt = 0;
    int last = (v0 > 0.0) - (v0 < 0.0);
    sgn_p = last;
    int sg = (v1 > 0.0) - (v1 < 0.0);
    if (sg != 0) { if (last != 0 && sg != last) ++cnt; last = sg; }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 1; k < 10; ++k) {
        const double v2 = cs[k] * ((ca[k] * x + cb[k]) * v1 - v0);
        v0 = v1; v1 = v2;
        sg = (v2 > 0.0) - (v2 < 0.0);
        if (sg != 0) { if (last != 0 && sg != last) ++cnt; last = sg; }
    }
    return cnt;
}

// One isolated root of polynomial h.  Exact roots (lo == hi) of the reversed polynomial must lie
// in 0 < |w| < 1: w = 0 is z = infinity, |w| = 1 belongs to p itself.
GEOM_HDI void emit_candidate(GroupShared& gs, int h, double lo, double hi) {
    if (h && lo == hi && !(fabs(lo) > 1e-12 && fabs(lo) < 1.0)) return;
    const int i = grp_atomic_inc(&gs.ncand[h]);
    if (i < 10) { gs.u.iso.cand_a[h][i] = lo; gs.u.iso.cand_b[h][i] = hi; }
}

// classify one interval (a,b] of polynomial h: emit a bracket, keep it pending, or drop it
GEOM_HDI void classify_interval(GroupShared& gs, int h, int nxt, double a, double b, int ca, int cb, int sa, int sb) {
    const int k = ca - cb;
    if (k <= 0) return;
    if (k == 1) {
        if (sb == 0) { emit_candidate(gs, h, b, b); return; }
        if (sa != 0 && sa != sb) { emit_candidate(gs, h, a, b); return; }
    }
    if (b - a < 5e-7) { emit_candidate(gs, h, 0.5 * (a + b), 0.5 * (a + b)); return; }      // unresolved cluster: report once
    const int i = grp_atomic_inc(&gs.npend[h][nxt]);
    if (i < kPendSlots) {
        PendInterval& q = gs.u.iso.pend[h][nxt][i];
        q.a = a; q.b = b; q.pk = pack_counts(ca, cb, sa, sb);
    }
}

// pts: x1[5][2] then x2[5][2] (20 doubles, identical in every lane).
// Eout[slot][9] / okout[slot]: solution of lane r (root index r in the ordered list).  Returns the
// number of root brackets (0 when the hypothesis is degenerate).  FINISH = false stops after the
// ordered bracket list (gs.v.fin, gs.nfinal): the caller finishes the roots itself (the CUDA kernel
// re-packs the roots of a whole CTA onto dense lanes).
// EXACT = false: a hypothesis that needs the exact root path is not solved here; -1 is returned and the
// caller queues it for a launch in which every group takes that path (no slow warp among fast ones).
template <bool FINISH, bool EXACT, class Exec>
GEOM_HD inline int five_point_group(Exec& ex, GroupShared& gs, const double* pts, double (*Eout)[9], bool* okout) {
    // ------------------------------------------------------------------ phase A: null space
    double a_[B200VO_GRP_SLOTS][5];
    int used_[B200VO_GRP_SLOTS], myrow_[B200VO_GRP_SLOTS];
    double n_[B200VO_GRP_SLOTS][4];

    ex.run([&](int r) {
        const int ls = lane_slot(r);
        if (r == 0) {
            for (int i = 0; i < 16; ++i) gs.key[i] = 0u;
            gs.fail = 0; gs.nfinal = 0;
            gs.ncand[0] = gs.ncand[1] = 0;
            for (int hh = 0; hh < 2; ++hh)
                for (int b = 0; b < 3; ++b) gs.npend[hh][b] = 0;
        }
        used_[ls] = (r >= 9); myrow_[ls] = -1;
        const int jr = (r < 9) ? r / 3 : 0, jc = (r < 9) ? r % 3 : 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 5; ++i) {
            const double xa = pts[2 * i], xb = pts[2 * i + 1], xc = pts[10 + 2 * i], xd = pts[10 + 2 * i + 1];
            const double u = (jr == 0) ? xc : ((jr == 1) ? xd : 1.0);
            const double v = (jc == 0) ? xa : ((jc == 1) ? xb : 1.0);
            a_[ls][i] = u * v;
        }
    });
    // complete pivoting: key of the largest remaining entry = magnitude | (7-row) << 4 | (15-column), i.e.
    // among equal magnitudes the lowest row, then the lowest column wins (the scan order of null_space_5x9;
    // synthetic forward motion has EXACTLY equal columns, and the choice decides which basis vector is N3)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int step = 0; step < 5; ++step) {
        ex.run([&](int r) {
            const int ls = lane_slot(r);
            if (!used_[ls]) {
                unsigned best = 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int i = step; i < 5; ++i) {
                    const unsigned k = mag_key(a_[ls][i]) | ((unsigned)(7 - i) << 4) | (unsigned)(15 - r);
                    best = (k > best) ? k : best;
                }
                grp_atomic_max(&gs.key[step], best);
            }
        });
        ex.run([&](int r) {
            const int ls = lane_slot(r);
            const unsigned k = gs.key[step];
            const int pr = 7 - (int)((k >> 4) & 7u), pcn = 15 - (int)(k & 15u);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = step + 1; i < 5; ++i)
                if (pr == i) { const double t = a_[ls][step]; a_[ls][step] = a_[ls][i]; a_[ls][i] = t; }
            if (r == pcn) {
                used_[ls] = 1; myrow_[ls] = step;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int i = 0; i < 5; ++i) gs.u.ex[i] = a_[ls][i];
            }
        });
        ex.run([&](int r) {
            const int ls = lane_slot(r);
            const double p = gs.u.ex[step];
            const double inv = (p != 0.0) ? 1.0 / p : 0.0;
            a_[ls][step] *= inv;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = 0; i < 5; ++i)
                if (i != step) a_[ls][i] -= gs.u.ex[i] * a_[ls][step];
        });
    }
    // null vectors: N[k][f_k] = 1, N[k][pivot column of row q] = -A[q][f_k]
    ex.run([&](int r) {
        const int ls = lane_slot(r);
        if (r < 9) gs.ibuf[r] = myrow_[ls];
    });
    int myk_[B200VO_GRP_SLOTS];
    ex.run([&](int r) {
        const int ls = lane_slot(r);
        int k = 0;
        for (int j = 0; j < 9; ++j) k += (j < r && gs.ibuf[j] < 0);
        myk_[ls] = k;
        if (r < 9 && myrow_[ls] < 0 && k < 4) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = 0; i < 5; ++i) gs.u.ex[8 + k * 5 + i] = a_[ls][i];
        }
    });
    ex.run([&](int r) {
        const int ls = lane_slot(r);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 4; ++k) {
            double v = 0.0;
            if (r < 9) {
                if (myrow_[ls] < 0) v = (k == myk_[ls]) ? 1.0 : 0.0;
                else v = -gs.u.ex[8 + k * 5 + myrow_[ls]];
            }
            n_[ls][k] = v;
        }
    });
    // Cholesky-QR, twice (orthonormal to working precision; Q equals the Gram-Schmidt basis)
    for (int pass = 0; pass < 2; ++pass) {
        ex.run([&](int r) {
            const int ls = lane_slot(r);
            if (r < 9) {
                int m = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int k1 = 0; k1 < 4; ++k1)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                    for (int k2 = k1; k2 < 4; ++k2) { gs.u.ex[30 + m * 9 + r] = n_[ls][k1] * n_[ls][k2]; ++m; }
            }
        });
        ex.run([&](int r) {
            double s = 0.0;
            for (int j = 0; j < 9; ++j) s += gs.u.ex[30 + r * 9 + j];
            gs.u.ex[r] = s;
        });
        ex.run([&](int r) {
            const int ls = lane_slot(r);
            // G index: (0,0)=0 (0,1)=1 (0,2)=2 (0,3)=3 (1,1)=4 (1,2)=5 (1,3)=6 (2,2)=7 (2,3)=8 (3,3)=9
            const double i0 = inv_sqrt(gs.u.ex[0]);
            const double l10 = gs.u.ex[1] * i0, l20 = gs.u.ex[2] * i0, l30 = gs.u.ex[3] * i0;
            const double i1 = inv_sqrt(gs.u.ex[4] - l10 * l10);
            const double l21 = (gs.u.ex[5] - l20 * l10) * i1, l31 = (gs.u.ex[6] - l30 * l10) * i1;
            const double i2 = inv_sqrt(gs.u.ex[7] - l20 * l20 - l21 * l21);
            const double l32 = (gs.u.ex[8] - l30 * l20 - l31 * l21) * i2;
            const double i3 = inv_sqrt(gs.u.ex[9] - l30 * l30 - l31 * l31 - l32 * l32);
            const double q0 = n_[ls][0] * i0;
            const double q1 = (n_[ls][1] - l10 * q0) * i1;
            const double q2 = (n_[ls][2] - l20 * q0 - l21 * q1) * i2;
            const double q3 = (n_[ls][3] - l30 * q0 - l31 * q1 - l32 * q2) * i3;
            n_[ls][0] = q0; n_[ls][1] = q1; n_[ls][2] = q2; n_[ls][3] = q3;
            if (pass == 1 && r < 9) {
                // A fixed generic rotation of the basis (orthogonal 4x4).  The solutions keep their E, only their
                // coordinates (x, y, z) in the basis change.  The elimination basis is aligned with the coordinate
                // axes of E, and on structured data (forward motion, integer pixel rows) several solutions then
                // share nearly the same z: clusters of near-multiple roots of det B(z).  Measured on the golden
                // pair: hypotheses needing the exact root path 7 % -> 0.5 %, samples with cv2's solution count
                // 98.0 % -> 99.7 %, cv2's accurate solutions reproduced 99.6 % -> 100 %, fewer bisection rounds.
                // (a random orthogonal matrix, numpy default_rng(123), 5th QR draw: among eight candidates it lost
                // the fewest true solutions on 6000 noise-free forward-motion / planar-scene samples -- 18 against
                // 153 without rotation -- structured matrices with equal entries keep part of the degeneracy)
                const double m00 = -0.78955596050694465, m01 = 0.51757844187046786, m02 = 0.21775985122655081, m03 = -0.2475774402741201;
                const double m10 = 0.16336327892265801, m11 = -0.18294721682977971, m12 = -0.07199833216814272, m13 = -0.96677763478417356;
                const double m20 = 0.52387011235234093, m21 = 0.83584503872302041, m22 = -0.15340523551034185, m23 = -0.058223795342407889;
                const double m30 = -0.2747251894376268, m31 = -0.0024379514532609733, m32 = -0.96117829827603407, m33 = 0.025620413820990812;
                gs.S.N[0][r] = m00 * q0 + m01 * q1 + m02 * q2 + m03 * q3;
                gs.S.N[1][r] = m10 * q0 + m11 * q1 + m12 * q2 + m13 * q3;
                gs.S.N[2][r] = m20 * q0 + m21 * q1 + m22 * q2 + m23 * q3;
                gs.S.N[3][r] = m30 * q0 + m31 * q1 + m32 * q2 + m33 * q3;
            }
        });
    }

    // ------------------------------------------------------------------ phase B: the ten cubic constraints
    // quads 0..5: Gram entries G_ik = sum_c E_ic E_kc, (i,k) = 00 01 02 11 12 22; quads 6..8: the
    // cofactors E_1b E_2c - E_1c E_2b.  Entry (i,c) of E is column 3i+c of N (a linear form in x,y,z,1).
    const double* Nb = &gs.S.N[0][0];
    ex.run([&](int r) {
        if (r < 9) {
            double q[10];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int t = 0; t < 10; ++t) q[t] = 0.0;
            int uc[3], vc[3];
            double w[3];
            if (r < 6) {
                const int i = (r >= 5) + (r >= 3);
                const int k = r - ((i == 0) ? 0 : ((i == 1) ? 2 : 3));
                for (int t = 0; t < 3; ++t) { uc[t] = 3 * i + t; vc[t] = 3 * k + t; w[t] = 1.0; }
            } else {
                const int p = r - 6, b = (p + 1) % 3, c = (p + 2) % 3;
                uc[0] = 3 + b; vc[0] = 6 + c; w[0] = 1.0;
                uc[1] = 3 + c; vc[1] = 6 + b; w[1] = -1.0;
                uc[2] = 0; vc[2] = 0; w[2] = 0.0;
            }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int t = 0; t < 3; ++t) lin_mul_acc(Nb + uc[t], Nb + vc[t], w[t], q);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int t = 0; t < 10; ++t) gs.u.ex[r * 10 + t] = q[t];
        }
    });
    double m_[B200VO_GRP_SLOTS][20];
    int pivcol_[B200VO_GRP_SLOTS];
    ex.run([&](int r) {
        const int ls = lane_slot(r);
        double acc[20];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 20; ++j) acc[j] = 0.0;
        const int i = (r >= 1) ? (r - 1) / 3 : 0, j = (r >= 1) ? (r - 1) % 3 : 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 3; ++k) {
            int qi, lc;
            bool sub;
            if (r == 0) { qi = 6 + k; lc = k; sub = false; }
            else {
                const int lo = (i < k) ? i : k, hi = (i < k) ? k : i;
                qi = (lo == 0) ? hi : ((lo == 1) ? 2 + hi : 5);
                lc = 3 * k + j;
                sub = (k == i);
            }
            double q[10];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int t = 0; t < 10; ++t) {
                double v = gs.u.ex[qi * 10 + t];
                if (sub) v -= 0.5 * (gs.u.ex[t] + gs.u.ex[30 + t] + gs.u.ex[50 + t]);       // - 1/2 tr(E E^T)
                q[t] = v;
            }
            quad_mul_acc(q, Nb + lc, acc);
        }
        double mx = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int jj = 0; jj < 20; ++jj) mx = fmax(mx, fabs(acc[jj]));
        const double inv = (mx > 0.0) ? 1.0 / mx : 1.0;                      // row equilibration
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int jj = 0; jj < 20; ++jj) m_[ls][jj] = acc[jj] * inv;
        pivcol_[ls] = -1;
        grp_atomic_max(&gs.key[5], mag_key(m_[ls][0]) | (unsigned)r);
    });

    // ------------------------------------------------------------------ phase C: Gauss-Jordan, lane = row
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int col = 0; col < 10; ++col) {
        ex.run([&](int r) {
            const int ls = lane_slot(r);
            const unsigned k = gs.key[5 + col];
            const int pl = (int)(k & 15u);
            if ((k >> 7) == 0u && r == 0) gs.fail = 1;                        // numerically zero column
            if (r == pl) {
                const double inv = 1.0 / m_[ls][col];
                pivcol_[ls] = col;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int j = col + 1; j < 20; ++j) { m_[ls][j] *= inv; gs.u.ex[j] = m_[ls][j]; }
            }
        });
        ex.run([&](int r) {
            const int ls = lane_slot(r);
            const int pl = (int)(gs.key[5 + col] & 15u);
            if (r != pl) {
                const double f = m_[ls][col];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int j = col + 1; j < 20; ++j) m_[ls][j] -= f * gs.u.ex[j];
            }
            if (col < 9 && pivcol_[ls] < 0) grp_atomic_max(&gs.key[5 + col + 1], mag_key(m_[ls][col + 1]) | (unsigned)r);
        });
    }

    // ------------------------------------------------------------------ phase D: B(z) and det B(z)
    ex.run([&](int r) {
        const int ls = lane_slot(r);
        if (pivcol_[ls] >= 4) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int t = 0; t < 10; ++t) gs.u.ex[(pivcol_[ls] - 4) * 10 + t] = m_[ls][10 + t];
        }
    });
    ex.run([&](int r) {
        if (r < 3) {
            const double* be = &gs.u.ex[(2 * r) * 10];
            const double* bf = &gs.u.ex[(2 * r + 1) * 10];
            double* Bx = gs.S.Bx[r]; double* By = gs.S.By[r]; double* B1 = gs.S.B1[r];
            Bx[3] = -bf[0]; Bx[2] = be[0] - bf[1]; Bx[1] = be[1] - bf[2]; Bx[0] = be[2];
            By[3] = -bf[3]; By[2] = be[3] - bf[4]; By[1] = be[4] - bf[5]; By[0] = be[5];
            B1[4] = -bf[6]; B1[3] = be[6] - bf[7]; B1[2] = be[7] - bf[8]; B1[1] = be[8] - bf[9]; B1[0] = be[9];
        }
    });
    ex.run([&](int r) {
        if (r < 3) {
            const int p = (r == 0) ? 1 : 0, q = (r == 2) ? 1 : 2;
            double mnr[7];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = 0; i < 7; ++i) mnr[i] = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = 0; i < 4; ++i)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int j = 0; j < 4; ++j) mnr[i + j] += gs.S.Bx[p][i] * gs.S.By[q][j] - gs.S.By[p][i] * gs.S.Bx[q][j];
            double part[11];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = 0; i < 11; ++i) part[i] = 0.0;
            const double sgn = (r == 1) ? -1.0 : 1.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = 0; i < 7; ++i)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int j = 0; j < 5; ++j) part[i + j] += sgn * mnr[i] * gs.S.B1[r][j];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = 0; i < 11; ++i) gs.u.ex[64 + r * 11 + i] = part[i];
        }
    });
    double cc_[B200VO_GRP_SLOTS][11];       // this lane's polynomial: lanes 0-4 p, lanes 5-9 the reversed p
    ex.run([&](int r) {
        const int ls = lane_slot(r);
        double c[11];
        double cmax = 0.0, csum = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 11; ++i) {
            c[i] = gs.u.ex[64 + i] + gs.u.ex[75 + i] + gs.u.ex[86 + i];
            cmax = fmax(cmax, fabs(c[i]));
            csum += c[i];
        }
        const bool good = (cmax > 0.0) && (cmax < 1e300) && (fabs(csum) < 1e300);      // also rejects NaN
        const double cinv = good ? 1.0 / cmax : 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 11; ++i) c[i] *= cinv;
        if (r == 0) {
            if (!good) gs.fail = 1;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = 0; i < 11; ++i) gs.S.c[i] = c[i];
        }
        const bool rev = (r >= 5);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 11; ++i) cc_[ls][i] = rev ? c[10 - i] : c[i];
    });
    if (ex.active && gs.fail) ex.active = false;          // degenerate sample: idle through the remaining phases

    // ------------------------------------------------------------------ phase E: root isolation
    // Half-group h (5 lanes) owns polynomial h.  Real roots of p in (-1,1] come from h = 0, those
    // with |z| > 1 from the reversed polynomial (h = 1, root w = 1/z in (-1,1)).
    ex.run([&](int r) {
        const int ls = lane_slot(r);
        const int h = r / 5, t = r - 5 * h;
        double ca[10], cb[10], cs[10];
        ca[0] = cb[0] = cs[0] = 0.0;
        bool ok = sturm_build10(cc_[ls], ca, cb, cs);
        ok = ok && (cc_[ls][10] != 0.0);
        if (t == 0) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int k = 0; k < 10; ++k) { gs.v.chain[h][0][k] = ca[k]; gs.v.chain[h][1][k] = cb[k]; gs.v.chain[h][2][k] = cs[k]; }
            gs.chain_ok[h] = ok ? 1 : 0;
        }
    });
    ex.run([&](int r) {
        const int ls = lane_slot(r);
        const int h = r / 5, t = r - 5 * h;
        if (gs.chain_ok[h]) {
            int sp;
            const int v = sturm_eval10(cc_[ls], gs.v.chain[h][0], gs.v.chain[h][1], gs.v.chain[h][2], -1.0 + 0.5 * t, sp);
            gs.ev_v[h][t] = v; gs.ev_s[h][t] = sp;
        }
    });
    ex.run([&](int r) {
        const int h = r / 5, t = r - 5 * h;
        if (gs.chain_ok[h]) {
            // parity check: the root count in (-1,1] is odd iff p changes sign between the ends; a
            // numerically unreliable chain (near-multiple roots) fails it -> exact fallback
            const int k = gs.ev_v[h][0] - gs.ev_v[h][4];
            const int sm = gs.ev_s[h][0], sp = gs.ev_s[h][4];
            const bool odd = (sm < 0) != (sp < 0);
            const bool bad = (k < 0) || sm == 0 || sp == 0 || (((k & 1) != 0) != odd);
            if (bad) { if (t == 0) gs.ibuf[10 + h] = 0; }
            else {
                if (t == 0) gs.ibuf[10 + h] = 1;
                if (t < 4) classify_interval(gs, h, 0, -1.0 + 0.5 * t, -1.0 + 0.5 * (t + 1), gs.ev_v[h][t], gs.ev_v[h][t + 1],
                                             gs.ev_s[h][t], gs.ev_s[h][t + 1]);
            }
        } else if (t == 0) gs.ibuf[10 + h] = 0;
    });
    // breadth-first bisection: buffer `cur` is read, `cur+1` filled, `cur+2` reset for the round after
    int cur = 0;
    for (int round = 0; round < 26; ++round) {
        if (!ex.any(ex.active && gs.npend[0][cur] + gs.npend[1][cur] > 0)) break;
        const int nxt = (cur + 1) % 3, clr = (cur + 2) % 3;
        ex.run([&](int r) {
            const int ls = lane_slot(r);
            const int h = r / 5, t = r - 5 * h;
            if (t == 0) gs.npend[h][clr] = 0;
            const int np = gs.npend[h][cur] < kPendSlots ? gs.npend[h][cur] : kPendSlots;
            for (int i = t; i < np; i += 5) {
                const PendInterval q = gs.u.iso.pend[h][cur][i];
                const int qca = q.pk & 0xFF, qcb = (q.pk >> 8) & 0xFF, qsa = ((q.pk >> 16) & 3) - 1, qsb = ((q.pk >> 18) & 3) - 1;
                const double mid = 0.5 * (q.a + q.b);
                int sm;
                const int cm = sturm_eval10(cc_[ls], gs.v.chain[h][0], gs.v.chain[h][1], gs.v.chain[h][2], mid, sm);
                classify_interval(gs, h, nxt, q.a, mid, qca, cm, qsa, sm);
                classify_interval(gs, h, nxt, mid, q.b, cm, qcb, sm, qsb);
            }
        });
        cur = nxt;
    }
    // Polynomials whose chain was unreliable (near-multiple roots, 2-6 % of the hypotheses) take the
    // exact path of real_roots_unit: the roots of the k-th derivative separate those of the (k-1)-th,
    // so level by level (linear derivative first) every interval between two consecutive roots of the
    // previous level holds at most one root, found by its sign change.  The intervals of one level are
    // independent: lane s takes interval s, ten levels are ten phases (the serial form needs ~55
    // bracket solves one after another).  Scaled derivatives T[k][i] = c[i+k] C(i+k,k) live in shared
    // memory (the work lists of the Sturm path are dead by now), so degrees can be run-time values.
    bool deferred = false;
    if (!EXACT && ex.active && (!gs.ibuf[10] || !gs.ibuf[11])) { deferred = true; ex.active = false; }
    for (int h = 0; EXACT && h < 2; ++h) {
        const bool need = ex.active && !gs.ibuf[10 + h];
        if (!ex.any(need)) continue;
        ex.run([&](int r) {
            if (!need) return;
            int n = 10;
            while (n > 0 && gs.S.c[h ? 10 - n : n] == 0.0) --n;
            if (r == 0) { gs.ibuf[0] = n; gs.ncand[h] = 0; }
            if (r < n) {                                   // row k = r: the k-th derivative / k!
                double* row = &gs.u.ex[r * 11 - r * (r - 1) / 2];
                for (int i = 0; i <= n - r; ++i) {
                    double binom = 1.0;
                    for (int j = 1; j <= r; ++j) binom = binom * (double)(i + j) / (double)j;
                    row[i] = gs.S.c[h ? 10 - (i + r) : i + r] * binom;
                }
            }
        });
        for (int k = 9; k >= 0; --k) {
            ex.run([&](int r) {
                if (!need) return;
                const int n = gs.ibuf[0];
                if (k > n - 1) return;
                const int dk = n - k, rd = (k + 1) & 1, wr = k & 1;
                const double* d = &gs.u.ex[k * 11 - k * (k - 1) / 2];
                double* val = &gs.u.ex[66];                // [2][10] roots of the current / previous level
                int cnt = 0;
                double lo = -1.0, hi = 1.0;
                if (k < n - 1)
                    for (int j = 0; j < 10; ++j)
                        if (gs.xflag[rd][j]) {
                            if (cnt == r - 1) lo = val[rd * 10 + j];
                            if (cnt == r) hi = val[rd * 10 + j];
                            ++cnt;
                        }
                bool has = false;
                double root = 0.0;
                if (r <= cnt) {                            // interval r of cnt + 1
                    double dc[11];                         // the level's polynomial in registers (zero above its degree)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                    for (int i = 0; i <= 10; ++i) dc[i] = (i <= dk) ? d[i] : 0.0;
                    const double flo = horner_reg10(dc, lo), fhi = horner_reg10(dc, hi);
                    if (flo == 0.0) { if (r == 0) { has = true; root = lo; } }
                    else if (fhi == 0.0) { has = true; root = hi; }
                    else if ((flo < 0.0) != (fhi < 0.0)) { has = true; root = bracket_root_fast_reg10(dc, lo, hi, flo); }
                }
                gs.xflag[wr][r] = has ? 1 : 0;
                if (has) {
                    val[wr * 10 + r] = root;
                    if (k == 0) emit_candidate(gs, h, root, root);
                }
            });
        }
    }
    // ordered list: roots of p ascending, then those of the reversed polynomial ascending in w
    // (the lists were filled through atomics: ranking makes the order deterministic)
    ex.run([&](int r) {
        const int n0 = gs.ncand[0] < 10 ? gs.ncand[0] : 10, n1 = gs.ncand[1] < 10 ? gs.ncand[1] : 10;
        for (int ci = r; ci < n0 + n1; ci += kGroupLanes) {
            const int h = (ci >= n0), me = ci - (h ? n0 : 0), n = h ? n1 : n0;
            const double a = gs.u.iso.cand_a[h][me], b = gs.u.iso.cand_b[h][me];
            int rank = 0;
            for (int i = 0; i < n; ++i) {
                const double ai = gs.u.iso.cand_a[h][i];
                rank += (ai < a || (ai == a && i < me));
            }
            const int pos = (h ? n0 : 0) + rank;
            if (pos < 10) { gs.v.fin.a[pos] = a; gs.v.fin.b[pos] = b; gs.v.fin.rev[pos] = h; }
        }
        if (r == 0) gs.nfinal = (n0 + n1 < 10) ? n0 + n1 : 10;
    });

    // ------------------------------------------------------------------ phase F: one root per lane
    const int nfinal = ex.active ? gs.nfinal : 0;
    if (FINISH) for (int ls = 0; ls < B200VO_GRP_SLOTS; ++ls) okout[ls] = false;
    if (FINISH) ex.run([&](int r) {
        const int ls = lane_slot(r);
        bool ok = false;
        if (r < nfinal) {
            RootBracket br;
            br.a = gs.v.fin.a[r]; br.b = gs.v.fin.b[r]; br.reversed = gs.v.fin.rev[r];
            ok = five_point_finish(gs.S, br, Eout[ls]);
        }
        okout[ls] = ok;
    });
    return deferred ? -1 : nfinal;
}

// host convenience (tests/hostsim): all solutions of one sample through the group algorithm
inline int five_point_group_host(const double x1[5][2], const double x2[5][2], double E[][9]) {
    static_assert(sizeof(GroupShared) % 16 == 0, "GroupShared must keep 16-byte alignment in an array");
    GroupShared gs;
    HostGroupExec ex;
    double pts[20];
    for (int i = 0; i < 5; ++i) { pts[2 * i] = x1[i][0]; pts[2 * i + 1] = x1[i][1]; pts[10 + 2 * i] = x2[i][0]; pts[10 + 2 * i + 1] = x2[i][1]; }
    double Eo[kGroupLanes][9];
    bool ok[kGroupLanes];
    const int nf = five_point_group<true, true>(ex, gs, pts, Eo, ok);
    int ns = 0;
    for (int r = 0; r < nf; ++r)
        if (ok[r]) { for (int k = 0; k < 9; ++k) E[ns][k] = Eo[r][k]; ++ns; }
    return ns;
}

}  // namespace geom
}  // namespace b200vo
