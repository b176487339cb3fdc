// TEST INFRASTRUCTURE ONLY: host build of the product's geometry header (geom_core.cuh) so the
// five-point solver / SVD / cheirality code can be checked against cv2 on the CPU-only box.
// The product library never links or loads this file.
#include <cstring>
#include "../../slam-visual-odometry_b200/csrc/geom_core.cuh"
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
