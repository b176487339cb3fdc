// TEST INFRASTRUCTURE ONLY: host build of the product's geometry header (geom_core.cuh) so the
// five-point solver / SVD / cheirality code can be checked against cv2 on the CPU-only box.
// The product library never links or loads this file.
#include <cstring>
#include "../../slam-visual-odometry_b200/csrc/geom_core.cuh"
#include "five_point_thread.h"
#include "../../slam-visual-odometry_b200/csrc/fivept_group.cuh"
using namespace b200vo::geom;
extern "C" {
int hs_five_point(const double* x1, const double* x2, double* E) {
    double a[5][2], b[5][2];
    std::memcpy(a, x1, sizeof(a)); std::memcpy(b, x2, sizeof(b));
    return five_point(a, b, reinterpret_cast<double (*)[9]>(E));
}
int hs_count_inliers(const double* E, const double* x1, const double* x2, int n, float thr2f, unsigned char* mask) {
    int c = 0;
    for (int i = 0; i < n; ++i) {
        const bool ok = sampson_inlier(E, x1[2 * i], x1[2 * i + 1], x2[2 * i], x2[2 * i + 1], thr2f);
        if (mask) mask[i] = ok;
        c += ok;
    }
    return c;
}
void hs_decompose(const double* E, double* R1, double* R2, double* t) { decompose_essential(E, R1, R2, t); }
int hs_cheirality(const double* R, const double* t, const double* x1, const double* x2, int n, double dist, unsigned char* ok) {
    int c = 0;
    for (int i = 0; i < n; ++i) { ok[i] = cheirality_ok(R, t, x1[2*i], x1[2*i+1], x2[2*i], x2[2*i+1], dist); c += ok[i]; }
    return c;
}
void hs_draw_samples(unsigned long long seed, long long pair_id, int nhyp, int m, int* out) {
    const uint64_t base = pair_stream(seed, pair_id);
    for (int h = 0; h < nhyp; ++h) draw_sample(base, h, m, out + 5 * h);
}
int hs_real_roots(const double* c, int deg, double* roots) { return real_roots_unit<10>(c, deg, roots); }
}
extern "C" int hs_real_roots_sturm(const double* c, int deg, double* roots) { return real_roots_sturm<10>(c, deg, roots); }
extern "C" int hs_count_inliers_fast(const double* E, const double* x1, const double* x2, int n, float thr2f, unsigned char* mask) {
    const SampsonThreshold t = make_sampson_threshold(thr2f);
    int c = 0;
    for (int i = 0; i < n; ++i) {
        const bool ok = sampson_inlier_fast(E, x1[2 * i], x1[2 * i + 1], x2[2 * i], x2[2 * i + 1], t);
        if (mask) mask[i] = ok;
        c += ok;
    }
    return c;
}
// the cooperative (10 lanes per hypothesis) solver, lanes emulated one after another
extern "C" int hs_five_point_group(const double* x1, const double* x2, double* E) {
    double a[5][2], b[5][2];
    std::memcpy(a, x1, sizeof(a)); std::memcpy(b, x2, sizeof(b));
    return five_point_group_host(a, b, reinterpret_cast<double (*)[9]>(E));
}
extern "C" int hs_group_shared_bytes() { return (int)sizeof(GroupShared); }
// debug: setup polynomial + brackets of both formulations
extern "C" int hs_dbg_thread(const double* x1, const double* x2, double* c, double* lo, double* hi, int* rev, double* N) {
    double a[5][2], b[5][2];
    std::memcpy(a, x1, sizeof(a)); std::memcpy(b, x2, sizeof(b));
    LocalMat M; FivePointSetup S;
    if (!five_point_setup(a, b, M, S)) return -1;
    std::memcpy(c, S.c, sizeof(S.c)); std::memcpy(N, S.N, sizeof(S.N));
    RootBracket br[10];
    const int nb = five_point_isolate(S.c, br);
    for (int i = 0; i < nb; ++i) { lo[i] = br[i].a; hi[i] = br[i].b; rev[i] = br[i].reversed; }
    return nb;
}
extern "C" int hs_dbg_group(const double* x1, const double* x2, double* c, double* lo, double* hi, int* rev, double* N, int* flags) {
    GroupShared gs; HostGroupExec ex;
    double pts[20];
    for (int i = 0; i < 5; ++i) { pts[2*i] = x1[2*i]; pts[2*i+1] = x1[2*i+1]; pts[10+2*i] = x2[2*i]; pts[10+2*i+1] = x2[2*i+1]; }
    double Eo[kGroupLanes][9]; bool ok[kGroupLanes];
    const int nf = five_point_group<true, true>(ex, gs, pts, Eo, ok);
    std::memcpy(c, gs.S.c, sizeof(gs.S.c)); std::memcpy(N, gs.S.N, sizeof(gs.S.N));
    for (int i = 0; i < nf; ++i) { lo[i] = gs.v.fin.a[i]; hi[i] = gs.v.fin.b[i]; rev[i] = gs.v.fin.rev[i]; }
    flags[0] = gs.ibuf[10]; flags[1] = gs.ibuf[11]; flags[2] = gs.chain_ok[0]; flags[3] = gs.chain_ok[1];
    { GroupShared g2; HostGroupExec e2; double E2[kGroupLanes][9]; bool o2[kGroupLanes]; flags[5] = five_point_group<true, false>(e2, g2, pts, E2, o2); }
    int okc = 0; for (int i = 0; i < nf; ++i) okc += ok[i];
    flags[4] = okc;
    return nf;
}
// 1 per point when the adjugate iteration of the DLT null vector converged (0: one-sided Jacobi path)
extern "C" int hs_cheirality_fast_path(const double* R, const double* t, const double* x1, const double* x2, int n, unsigned char* fast) {
    int c = 0;
    for (int i = 0; i < n; ++i) { fast[i] = (unsigned char)cheirality_fast_path(R, t, x1[2*i], x1[2*i+1], x2[2*i], x2[2*i+1]); c += fast[i]; }
    return c;
}
// register-resident Horner / bracket solver of the exact root path vs the run-time-degree versions: must agree bit for bit
// (leading zero coefficients leave every intermediate value unchanged)
extern "C" int hs_bracket_root_both(const double* c, int deg, double a, double b, double* out_runtime, double* out_reg10) {
    double cr[11];
    for (int i = 0; i <= 10; ++i) cr[i] = (i <= deg) ? c[i] : 0.0;
    const double fa = horner(c, deg, a), fa2 = horner_reg10(cr, a);
    const double fb = horner(c, deg, b), fb2 = horner_reg10(cr, b);
    if (std::memcmp(&fa, &fa2, 8) != 0 || std::memcmp(&fb, &fb2, 8) != 0) return -1;
    if (fa == 0.0 || fb == 0.0 || (fa < 0.0) == (fb < 0.0)) return 0;
    *out_runtime = bracket_root_fast(c, deg, a, b, fa);
    *out_reg10 = bracket_root_fast_reg10(cr, a, b, fa2);
    return 1;
}
