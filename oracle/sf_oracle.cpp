/*
 * sf_oracle.cpp — CPU oracle for the StaticFusion joint odometry + segmentation solver.
 *
 * TEST INFRASTRUCTURE ONLY (see sf_oracle.h).  Pinned against the reference's own sources compiled
 * with a header shim (oracle/ref_shim): the reference-literal policy below reproduces them bit for
 * bit (tests/test_oracle_vs_reference.py).  This file is a restatement of
 *   FrontEnd.cpp:256-892,1071-1146   SegmentationBackground.cpp:53-197   KMeans.cpp:52-391
 * Each function cites the reference lines it follows.  Images are stored row-major here
 * but every loop keeps the reference's traversal order (u outer, v inner = Eigen
 * column-major order) because that order defines the float summation order.
 *
 * Three accumulation policies (sf_oracle.h):
 *   ORC_ACCUM_F32   reference-literal sequential float sums, float small algebra.
 *   ORC_ACCUM_EXACT every cross-pixel sum is an order-independent fixed-point (integer) sum: k-means centres,
 *                   warp splat, seg prior, mean|B|, per-label residuals, and the normal equations / |res|^2
 *                   (products rounded to 2^-40 of the column-bound product); double small algebra.
 *                   This is the numerics contract of the CUDA path.
 *   ORC_ACCUM_F64   like EXACT but plain double sums for the normal equations / |res|^2: the yardstick.
 *
 * Third-party arithmetic that is NOT in the reference tree (Eigen LDLT / inverse /
 * SelfAdjointEigenSolver / colPivHouseholderQr / matrix exp+log, MRPT CPose3D) is
 * replaced by documented closed forms: unpivoted LDL^T with zero-pivot handling,
 * cyclic Jacobi, Rodrigues exp/log.  Build with -ffp-contract=off.
 */
#include "sf_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <utility>
#include <vector>

namespace {

constexpr int NC = ORC_NUM_CLUSTERS;

struct Img {
    int rows = 0, cols = 0;
    std::vector<float> a;
    void resize(int r, int c) { rows = r; cols = c; a.assign((size_t)r * c, 0.f); }
    void fill(float v) { std::fill(a.begin(), a.end(), v); }
    float& operator()(int v, int u) { return a[(size_t)v * cols + u]; }
    float operator()(int v, int u) const { return a[(size_t)v * cols + u]; }
};
struct ImgI {
    int rows = 0, cols = 0;
    std::vector<int32_t> a;
    void resize(int r, int c) { rows = r; cols = c; a.assign((size_t)r * c, 0); }
    void fill(int v) { std::fill(a.begin(), a.end(), v); }
    int32_t& operator()(int v, int u) { return a[(size_t)v * cols + u]; }
    int32_t operator()(int v, int u) const { return a[(size_t)v * cols + u]; }
};

inline float sq(float x) { return x * x; }

/* fixed-point quantisation used by the EXACT policy: round-to-nearest-even of x*2^s */
inline int64_t fixq(float x, int s) { return (int64_t)llrintf(ldexpf(x, s)); }
inline double fixval(int64_t acc, int s) { return std::ldexp((double)acc, -s); }

/* EXACT policy for the normal equations and |res|^2 (mirrors sf_device.cuh): every column of the weighted system is
 * scaled by a power of two so that its magnitude stays below 2^10; each product of two scaled entries (exact in double:
 * 24 x 24 bits) is rounded ONCE to the nearest multiple of 2^-20 (round-half-even) and those fixed-point values are summed
 * exactly as integers: associative, hence independent of the reduction order.  On the GPU the rounding and the addition
 * are one double-precision FMA onto an accumulator that stays inside the binade of 1.5 * 2^32 (its ulp is 2^-20 there, so
 * fma(a, b, acc) == acc + round_to_2^-20(a * b) exactly); here it is nearbyint + an integer sum.  The quantum is 2^-40 of
 * (column bound x column bound), which puts the sums ~1e-9 (relative) from plain double sums at every pyramid level
 * (scripts/sweep_numerics.py; the first-round policy rounded each product to 2^-20 of that bound). */
constexpr int ORC_QSCALE_BITS = 10; /* scaled columns stay below 2^10 */
constexpr int ORC_QFRAC_BITS = 20;  /* products are rounded to multiples of 2^-20 */
constexpr int ORC_LABEL_BITS = 30;  /* per-label residual terms: round(x * 2^s) < 2^30 (cvt.rni.s32.f32 on the GPU) */
inline int scale_exponent(float bound) {
    if (!(bound > 0.f) || !std::isfinite(bound)) return 0;
    int e;
    (void)frexpf(bound, &e); /* bound = m * 2^e, m in [0.5, 1) */
    int s = ORC_QSCALE_BITS - e;
    if (s > 100) s = 100;
    if (s < -100) s = -100;
    return s;
}
inline int64_t qprod(float a, float b) { return (int64_t)std::nearbyint(std::ldexp((double)a * (double)b, ORC_QFRAC_BITS)); }

struct StageTimer { /* adds the scope's wall time to one stage cell (nested scopes subtract themselves from the outer one) */
    double* cell; double* outer;
    std::chrono::steady_clock::time_point t0;
    StageTimer(double* c, double* o = nullptr) : cell(c), outer(o), t0(std::chrono::steady_clock::now()) {}
    ~StageTimer() {
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        *cell += dt;
        if (outer) *outer -= dt;
    }
};

/* =====================================================================================
 * Small dense algebra (templated on float / double)
 * ===================================================================================== */

/* Unpivoted LDL^T of a symmetric n x n matrix (row-major, full storage).  Pivots <= tiny are
 * treated as exactly singular: the column of L is zeroed and the matching solution
 * component is 0 (stands in for Eigen's pivoted LDLT zero-pivot rule at
 * FrontEnd.cpp:642 / SegmentationBackground.cpp:168).  Returns number of zero pivots. */
template <class T>
int ldlt_factor(int n, T* A, unsigned char* zero) {
    const T tiny = (T)1e-20;
    int nz = 0;
    for (int j = 0; j < n; j++) {
        T d = A[j * n + j];
        for (int k = 0; k < j; k++) d -= (A[j * n + k] * A[j * n + k]) * A[k * n + k];
        A[j * n + j] = d;
        if (!(d > tiny)) {
            zero[j] = 1; nz++;
            for (int i = j + 1; i < n; i++) A[i * n + j] = (T)0;
            continue;
        }
        zero[j] = 0;
        for (int i = j + 1; i < n; i++) {
            T s = A[i * n + j];
            for (int k = 0; k < j; k++) s -= (A[i * n + k] * A[j * n + k]) * A[k * n + k];
            A[i * n + j] = s / d;
        }
    }
    return nz;
}
template <class T>
void ldlt_solve_factored(int n, const T* A, const unsigned char* zero, const T* b, T* x) {
    for (int i = 0; i < n; i++) {
        T s = b[i];
        for (int k = 0; k < i; k++) s -= A[i * n + k] * x[k];
        x[i] = s;
    }
    for (int i = 0; i < n; i++) x[i] = zero[i] ? (T)0 : x[i] / A[i * n + i];
    for (int i = n - 1; i >= 0; i--) { /* descending k: the order a column sweep produces (x_k is used as soon as it is final) */
        T s = x[i];
        for (int k = n - 1; k > i; k--) s -= A[k * n + i] * x[k];
        x[i] = s;
    }
}

/* Gaussian elimination with partial pivoting on [A | B] (row-major n x n, n x m): the reference-literal policy's stand-in
 * for Eigen's general `inverse()` (FrontEnd.cpp:689,755,800) and `colPivHouseholderQr().solve()` (:729,741,755). */
template <class T>
bool gauss_solve(int n, int m, T* A, T* B) {
    for (int c = 0; c < n; c++) {
        int p = c;
        for (int r = c + 1; r < n; r++) if (std::fabs(A[r * n + c]) > std::fabs(A[p * n + c])) p = r;
        if (A[p * n + c] == (T)0) return false;
        if (p != c) {
            for (int k = 0; k < n; k++) std::swap(A[p * n + k], A[c * n + k]);
            for (int k = 0; k < m; k++) std::swap(B[p * m + k], B[c * m + k]);
        }
        for (int r = c + 1; r < n; r++) {
            const T f = A[r * n + c] / A[c * n + c];
            if (f == (T)0) continue;
            for (int k = c; k < n; k++) A[r * n + k] -= f * A[c * n + k];
            for (int k = 0; k < m; k++) B[r * m + k] -= f * B[c * m + k];
        }
    }
    for (int k = 0; k < m; k++)
        for (int r = n - 1; r >= 0; r--) {
            T s = B[r * m + k];
            for (int c = r + 1; c < n; c++) s -= A[r * n + c] * B[c * m + k];
            B[r * m + k] = s / A[r * n + r];
        }
    return true;
}
template <class T>
void general_inverse(int n, const T* A, T* inv) {
    std::vector<T> M(A, A + (size_t)n * n), B((size_t)n * n, (T)0);
    for (int i = 0; i < n; i++) B[(size_t)i * n + i] = (T)1;
    const bool ok = gauss_solve<T>(n, n, M.data(), B.data());
    for (int i = 0; i < n * n; i++) inv[i] = ok ? B[i] : std::numeric_limits<T>::quiet_NaN();
}

/* Cyclic Jacobi for a symmetric 6x6 (stands in for SelfAdjointEigenSolver, FrontEnd.cpp:719).
 * V columns are eigenvectors.  No sorting: the motion filter is invariant to the order. */
template <class T>
void jacobi_eig6(const T* Ain, T* ev, T* V) {
    const int n = 6;
    T A[36];
    for (int i = 0; i < 36; i++) A[i] = Ain[i];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) V[i * n + j] = (i == j) ? (T)1 : (T)0;
    for (int sweep = 0; sweep < 30; sweep++) {
        T off = 0, diag = 0;
        for (int p = 0; p < n; p++) {
            diag += A[p * n + p] * A[p * n + p];
            for (int q = p + 1; q < n; q++) off += A[p * n + q] * A[p * n + q];
        }
        if (!(off > (T)1e-32 * diag)) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                const T apq = A[p * n + q];
                if (apq == (T)0) continue;
                const T theta = (A[q * n + q] - A[p * n + p]) / ((T)2 * apq);
                const T t = (theta >= (T)0 ? (T)1 : (T)-1) / (std::fabs(theta) + std::sqrt(theta * theta + (T)1));
                const T c = (T)1 / std::sqrt(t * t + (T)1);
                const T s = t * c;
                for (int k = 0; k < n; k++) {
                    const T akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    const T apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    const T vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; i++) ev[i] = A[i * n + i];
}

/* SE(3) exponential of twist (t, w) -> 4x4 row-major (stands in for Matrix4f::exp(), FrontEnd.cpp:766).
 * Series below theta = 0.05 so that the common small-motion case uses only + - * /. */
template <class T>
void se3_exp(const T* xi, T* M) {
    const T wx = xi[3], wy = xi[4], wz = xi[5];
    const T th2 = wx * wx + wy * wy + wz * wz;
    T A, B, C;
    if (th2 < (T)0.0025) {
        A = (T)1 + th2 * ((T)-1 / 6 + th2 * ((T)1 / 120 + th2 * ((T)-1 / 5040 + th2 * ((T)1 / 362880))));
        B = (T)0.5 + th2 * ((T)-1 / 24 + th2 * ((T)1 / 720 + th2 * ((T)-1 / 40320 + th2 * ((T)1 / 3628800))));
        C = (T)1 / 6 + th2 * ((T)-1 / 120 + th2 * ((T)1 / 5040 + th2 * ((T)-1 / 362880 + th2 * ((T)1 / 39916800))));
    } else {
        const T th = std::sqrt(th2);
        const T sh = std::sin((T)0.5 * th);
        A = std::sin(th) / th;
        B = (T)2 * sh * sh / th2;
        C = (th - std::sin(th)) / (th2 * th);
    }
    const T K[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    T K2[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            T s = 0;
            for (int k = 0; k < 3; k++) s += K[i * 3 + k] * K[k * 3 + j];
            K2[i * 3 + j] = s;
        }
    T R[9], Vm[9];
    for (int i = 0; i < 9; i++) {
        const T I = (i % 4 == 0) ? (T)1 : (T)0;
        R[i] = I + A * K[i] + B * K2[i];
        Vm[i] = I + B * K[i] + C * K2[i];
    }
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) M[i * 4 + j] = R[i * 3 + j];
        M[i * 4 + 3] = Vm[i * 3 + 0] * xi[0] + Vm[i * 3 + 1] * xi[1] + Vm[i * 3 + 2] * xi[2];
    }
    M[12] = 0; M[13] = 0; M[14] = 0; M[15] = 1;
}

/* SE(3) logarithm -> twist (stands in for Matrix4f::log(), FrontEnd.cpp:736,769). */
template <class T>
void se3_log(const T* M, T* xi) {
    const T sx = (T)0.5 * (M[2 * 4 + 1] - M[1 * 4 + 2]);
    const T sy = (T)0.5 * (M[0 * 4 + 2] - M[2 * 4 + 0]);
    const T sz = (T)0.5 * (M[1 * 4 + 0] - M[0 * 4 + 1]);
    const T s2 = sx * sx + sy * sy + sz * sz;
    const T c = (T)0.5 * (M[0] + M[5] + M[10] - (T)1);
    T fac, th2, D;
    if (s2 < (T)0.0025 && c > (T)0) {
        /* theta/sin(theta) as a series in s^2 = sin^2(theta): asin(s)/s */
        fac = (T)1 + s2 * ((T)1 / 6 + s2 * ((T)3 / 40 + s2 * ((T)15 / 336 + s2 * ((T)105 / 3456 + s2 * ((T)945 / 42240)))));
        th2 = s2 * fac * fac;
        D = (T)1 / 12 + th2 * ((T)1 / 720 + th2 * ((T)1 / 30240 + th2 * ((T)1 / 1209600)));
    } else {
        const T s = std::sqrt(s2);
        const T th = std::atan2(s, c);
        fac = (s > (T)0) ? th / s : (T)1;
        th2 = th * th;
        const T sh = std::sin((T)0.5 * th);
        const T Bc = (T)2 * sh * sh / th2;
        const T Ac = std::sin(th) / th;
        D = ((T)1 - Ac / ((T)2 * Bc)) / th2;
    }
    const T wx = fac * sx, wy = fac * sy, wz = fac * sz;
    const T K[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    T K2[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            T s = 0;
            for (int k = 0; k < 3; k++) s += K[i * 3 + k] * K[k * 3 + j];
            K2[i * 3 + j] = s;
        }
    for (int i = 0; i < 3; i++) {
        T acc = 0;
        for (int j = 0; j < 3; j++) {
            const T I = (i == j) ? (T)1 : (T)0;
            acc += (I - (T)0.5 * K[i * 3 + j] + D * K2[i * 3 + j]) * M[j * 4 + 3];
        }
        xi[i] = acc;
    }
    xi[3] = wx; xi[4] = wy; xi[5] = wz;
}

/* Rigid inverse of a 4x4 (stands in for the fixed-size Matrix4f::inverse(), FrontEnd.cpp:800). */
template <class T>
void rigid_inverse(const float* M, float* out) {
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) out[i * 4 + j] = M[j * 4 + i];
        T s = 0;
        for (int j = 0; j < 3; j++) s += (T)M[j * 4 + i] * (T)M[j * 4 + 3];
        out[i * 4 + 3] = (float)(-s);
    }
    out[12] = 0; out[13] = 0; out[14] = 0; out[15] = 1;
}

}  // namespace

/* =====================================================================================
 * Oracle state
 * ===================================================================================== */
struct orc_ctx {
    orc_params p;
    int accum;
    int rows, cols;
    int pyr_levels;

    std::vector<Img> intensityPyr, intensityPredPyr, intensityInterPyr, intensityWarpedPyr;
    std::vector<Img> depthPyr, depthPredPyr, depthInterPyr, depthWarpedPyr;
    std::vector<Img> xxPyr, xxInterPyr, xxPredPyr, xxWarpedPyr;
    std::vector<Img> yyPyr, yyInterPyr, yyPredPyr, yyWarpedPyr;
    Img depthCurrent, intensityCurrent, depthPrediction, intensityPrediction;
    Img dcu, dcv, dct, ddu, ddv, ddt, weights_c, weights_d;
    std::vector<unsigned char> Null;  /* row-major rows_i x cols_i of the active level */
    float convMask[16];

    float T_odometry[16];
    float twist_odometry[6], twist_level_odometry[6], twist_odometry_old[6];
    double est_cov[36];

    int rows_i, cols_i, image_level, level;
    std::vector<std::pair<int, int>> validPixels;

    std::vector<ImgI> clusterAllocation;
    float kmeans[3][NC];
    bool connectivity[NC][NC];

    float b_segm[NC], b_prior[NC], lambda_t_w[NC];
    Img b_segm_perpixel;
    float perClusterAverageResidual[NC];

    /* 5-frame history (StaticFusion.h:92-99); the drivers write the ring after every frame (StaticFusion-datasets.cpp:182-184) */
    static constexpr int bufferLength = 5;
    Img depthBuffer[bufferLength], intensityBuffer[bufferLength];
    float odomBuffer[bufferLength][16];
    Img depthWarpedRefference, intensityWarpedRefference, cumulativeResiduals;

    float max_wc_raw, max_wd_raw;
    int status, total_irls;
    /* per-stage wall time (std::chrono::steady_clock): pyramid, k-means, warp, linearise, IRLS,
     * seg-solve, pose update, per-pixel image */
    double stage_s[ORC_NUM_STAGES] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<float> trace;
    float* cur_trace;

    void init(const orc_params& pp, int accum_mode);
    void levelDims(int L, int& r, int& c) const { r = rows >> L; c = cols >> L; }

    void createImagePyramid(bool old_im);
    void initializeKMeans();
    void kMeans3DCoord();
    void computeRegionConnectivity();
    void createClustersPyramidUsingKMeans();
    void warpImagesAccurateInverse();
    void calculateCoord();
    void calculateDerivatives();
    void computeWeights();
    void computeSegPrior();
    void solveOdometryAndSegmJoint();
    void solveSegmIteration(const float* aver_res, float aver_res_overall, const double* lap);
    void filterEstimateAndComputeT(float* twist);
    void runSolver(bool create_image_pyr, int stop_step);
    void buildSegmImage();
    void computeResidualsAgainstPreviousImage(int index);
};

void orc_ctx::init(const orc_params& pp, int accum_mode) {
    p = pp;
    accum = accum_mode;
    rows = p.rows; cols = p.cols;
    pyr_levels = p.ctf_levels; /* width == cols so round(log2(width/cols)) == 0, FrontEnd.cpp:108 */
    auto rs = [&](std::vector<Img>& v) { v.resize(pyr_levels); for (int i = 0; i < pyr_levels; i++) v[i].resize(rows >> i, cols >> i); };
    rs(intensityPyr); rs(intensityPredPyr); rs(intensityInterPyr); rs(intensityWarpedPyr);
    rs(depthPyr); rs(depthPredPyr); rs(depthInterPyr); rs(depthWarpedPyr);
    rs(xxPyr); rs(xxInterPyr); rs(xxPredPyr); rs(xxWarpedPyr);
    rs(yyPyr); rs(yyInterPyr); rs(yyPredPyr); rs(yyWarpedPyr);
    depthCurrent.resize(rows, cols); intensityCurrent.resize(rows, cols);
    depthPrediction.resize(rows, cols); intensityPrediction.resize(rows, cols);
    dcu.resize(rows, cols); dcv.resize(rows, cols); dct.resize(rows, cols);
    ddu.resize(rows, cols); ddv.resize(rows, cols); ddt.resize(rows, cols);
    weights_c.resize(rows, cols); weights_d.resize(rows, cols);
    Null.assign((size_t)rows * cols, 0);
    clusterAllocation.resize(pyr_levels);
    for (int i = 0; i < pyr_levels; i++) clusterAllocation[i].resize(rows >> i, cols >> i);
    /* FrontEnd.cpp:146-149 */
    const float v_mask[4] = {1.f, 2.f, 2.f, 1.f};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) convMask[i + 4 * j] = v_mask[i] * v_mask[j] / 36.f;
    b_segm_perpixel.resize(rows, cols);
    b_segm_perpixel.fill(0.5f);
    for (int l = 0; l < NC; l++) {
        b_segm[l] = 0.5f; b_prior[l] = 0.f; lambda_t_w[l] = 0.f;
        perClusterAverageResidual[l] = std::numeric_limits<float>::quiet_NaN(); /* FrontEnd.cpp:105 */
        for (int k = 0; k < 3; k++) kmeans[k][l] = 0.f;
        for (int m = 0; m < NC; m++) connectivity[l][m] = (l == m);
    }
    for (int b = 0; b < bufferLength; b++) { /* FrontEnd.cpp:96-103 (odomBuffer is left uninitialised there) */
        depthBuffer[b].resize(rows, cols); intensityBuffer[b].resize(rows, cols);
        for (int i = 0; i < 16; i++) odomBuffer[b][i] = (i % 5 == 0) ? 1.f : 0.f;
    }
    depthWarpedRefference.resize(rows, cols); intensityWarpedRefference.resize(rows, cols); cumulativeResiduals.resize(rows, cols);
    for (int i = 0; i < 16; i++) T_odometry[i] = (i % 5 == 0) ? 1.f : 0.f;
    for (int i = 0; i < 6; i++) twist_odometry[i] = twist_level_odometry[i] = twist_odometry_old[i] = 0.f;
    for (int i = 0; i < 36; i++) est_cov[i] = 0;
    status = 0; total_irls = 0;
    max_wc_raw = max_wd_raw = 0.f;
    trace.assign((size_t)p.ctf_levels * p.max_iter_per_level * ORC_TRACE_STEP, 0.f);
    cur_trace = nullptr;
}

/* FrontEnd.cpp:256-391 */
void orc_ctx::createImagePyramid(bool old_im) {
    StageTimer tm(&stage_s[0]);
    const float max_depth_dif = 0.1f;
    for (int i = 0; i < pyr_levels; i++) {
        const int ci = cols >> i, ri = rows >> i;
        Img& depth_here = old_im ? depthPredPyr[i] : depthPyr[i];
        Img& intensity_here = old_im ? intensityPredPyr[i] : intensityPyr[i];
        Img& xx_here = old_im ? xxPredPyr[i] : xxPyr[i];
        Img& yy_here = old_im ? yyPredPyr[i] : yyPyr[i];
        if (i == 0) { /* :282-292 */
            depth_here.a = old_im ? depthPrediction.a : depthCurrent.a;
            intensity_here.a = old_im ? intensityPrediction.a : intensityCurrent.a;
        } else {
            const Img& depth_prev = old_im ? depthPredPyr[i - 1] : depthPyr[i - 1];
            const Img& intensity_prev = old_im ? intensityPredPyr[i - 1] : intensityPyr[i - 1];
            for (int u = 0; u < ci; u++)
                for (int v = 0; v < ri; v++) {
                    const int u2 = 2 * u, v2 = 2 * v;
                    if ((v > 0) && (v < ri - 1) && (u > 0) && (u < ci - 1)) { /* :305 inner pixels */
                        float db[16], ib[16]; /* column-major 4x4 block at (v2-1,u2-1), :308-309 */
                        for (int c = 0; c < 4; c++)
                            for (int r = 0; r < 4; r++) {
                                db[r + 4 * c] = depth_prev(v2 - 1 + r, u2 - 1 + c);
                                ib[r + 4 * c] = intensity_prev(v2 - 1 + r, u2 - 1 + c);
                            }
                        float depths[4] = {db[5], db[6], db[9], db[10]}; /* :311 */
                        if (depths[1] < depths[0]) std::swap(depths[1], depths[0]);
                        if (depths[3] < depths[2]) std::swap(depths[3], depths[2]);
                        const float dcenter = (depths[3] < depths[1]) ? std::max(depths[3], depths[0]) : std::max(depths[1], depths[2]);
                        if (dcenter != 0.f) {
                            float sum_d = 0.f, sum_c = 0.f, weight = 0.f;
                            for (int k = 0; k < 16; k++) { /* :323-333 */
                                const float abs_dif = std::fabs(db[k] - dcenter);
                                if (abs_dif < max_depth_dif) {
                                    const float aux_w = convMask[k] * (max_depth_dif - abs_dif);
                                    weight += aux_w;
                                    sum_d += aux_w * db[k];
                                    sum_c += aux_w * ib[k];
                                }
                            }
                            depth_here(v, u) = sum_d / weight;
                            intensity_here(v, u) = sum_c / weight;
                        } else { /* :339-343; Eigen's packet reduction order is not reproducible: sequential k */
                            float s = 0.f;
                            for (int k = 0; k < 16; k++) s += convMask[k] * ib[k];
                            intensity_here(v, u) = s;
                            depth_here(v, u) = 0.f;
                        }
                    } else { /* :347-373 boundary, 2x2 block in column-major order */
                        const float d4[4] = {depth_prev(v2, u2), depth_prev(v2 + 1, u2), depth_prev(v2, u2 + 1), depth_prev(v2 + 1, u2 + 1)};
                        const float i4[4] = {intensity_prev(v2, u2), intensity_prev(v2 + 1, u2), intensity_prev(v2, u2 + 1), intensity_prev(v2 + 1, u2 + 1)};
                        intensity_here(v, u) = 0.25f * (((i4[0] + i4[1]) + i4[2]) + i4[3]);
                        float new_d = 0.f; unsigned cont = 0;
                        for (int k = 0; k < 4; k++)
                            if (d4[k] != 0.f) { new_d += d4[k]; cont++; }
                        depth_here(v, u) = cont ? new_d / float(cont) : 0.f;
                    }
                }
        }
        /* :378-388 */
        const float inv_f_i = 2.f * std::tan(0.5f * p.fovh) / float(ci);
        const float disp_u_i = 0.5f * float(ci - 1);
        const float disp_v_i = 0.5f * float(ri - 1);
        for (int u = 0; u < ci; u++)
            for (int v = 0; v < ri; v++) {
                yy_here(v, u) = (inv_f_i * (float(v) - disp_v_i)) * depth_here(v, u);
                xx_here(v, u) = (inv_f_i * (float(u) - disp_u_i)) * depth_here(v, u);
            }
    }
}

struct IndexAndDistance { int idx; float distance; };
/* KMeans.cpp:57-60,182: std::sort by distance only (unstable); restated as stable (distance, index)
 * — differs from libstdc++ introsort only in the order of exactly-equal distances. */
static void sortDistances(IndexAndDistance* d) {
    std::stable_sort(d, d + NC, [](const IndexAndDistance& a, const IndexAndDistance& b) { return a.distance < b.distance; });
}
static inline float sqnorm3(const float* a, const float* b) {
    const float d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    return (d0 * d0 + d1 * d1) + d2 * d2;
}

/* KMeans.cpp:63-135 */
void orc_ctx::initializeKMeans() {
    const int rows_km = rows / 2, cols_km = cols / 2;
    const Img& depth_ref = depthPyr[1];
    ImgI& labels_ref = clusterAllocation[1];
    labels_ref.fill(NC);
    unsigned u_label[NC], v_label[NC];
    const unsigned vert_div = (unsigned)std::ceil(std::sqrt((double)NC));
    const float u_div = float(cols_km) / float(NC + 1);
    const float v_div = float(rows_km) / float(vert_div + 1);
    for (unsigned i = 0; i < (unsigned)NC; i++) {
        u_label[i] = (unsigned)std::round((i + 1) * u_div);
        v_label[i] = (unsigned)std::round((i % vert_div + 1) * v_div);
    }
    std::vector<float> depth_sorted[NC];
    for (int u = 0; u < cols_km; u++)
        for (int v = 0; v < rows_km; v++)
            if (depth_ref(v, u) != 0.f) {
                unsigned min_dist = 1000000, quad_dist; /* :91 */
                unsigned ini_label = NC;
                for (unsigned l = 0; l < (unsigned)NC; l++) {
                    const int dv = v - (int)v_label[l], du = u - (int)u_label[l];
                    quad_dist = (unsigned)(dv * dv + du * du);
                    if (quad_dist < min_dist) { ini_label = l; min_dist = quad_dist; }
                }
                labels_ref(v, u) = (int)ini_label;
                if (ini_label < (unsigned)NC) depth_sorted[ini_label].push_back(depth_ref(v, u));
            }
    const float inv_f_i = 2.f * std::tan(0.5f * p.fovh) / float(cols_km);
    const float disp_u_i = 0.5f * float(cols_km - 1);
    const float disp_v_i = 0.5f * float(rows_km - 1);
    for (int l = 0; l < NC; l++) {
        const size_t size_label = depth_sorted[l].size();
        const size_t med_pos = size_label / 2;
        if (size_label > 0) {
            std::nth_element(depth_sorted[l].begin(), depth_sorted[l].begin() + med_pos, depth_sorted[l].end());
            kmeans[0][l] = depth_sorted[l][med_pos];
            kmeans[1][l] = (float(u_label[l]) - disp_u_i) * kmeans[0][l] * inv_f_i;
            kmeans[2][l] = (float(v_label[l]) - disp_v_i) * kmeans[0][l] * inv_f_i;
        } else {
            kmeans[0][l] = kmeans[1][l] = kmeans[2][l] = 0.f;
        }
    }
}

/* KMeans.cpp:137-295 */
void orc_ctx::kMeans3DCoord() {
    const int rows_km = rows / 2, cols_km = cols / 2;
    const int iter_kmeans = 10;
    const Img& depth_ref = depthPyr[1];
    const Img& xx_ref = xxPyr[1];
    const Img& yy_ref = yyPyr[1];
    ImgI& labels_lowres = clusterAllocation[1];
    initializeKMeans();

    IndexAndDistance cluster_distances[NC][NC];
    float centers_a[NC][3], centers_b[NC][3];
    int count[NC];
    for (int c = 0; c < NC; c++)
        for (int r = 0; r < 3; r++) centers_a[c][r] = kmeans[r][c];

    for (int it = 0; it < iter_kmeans - 1; it++) {
        int64_t fix_b[NC][3];
        for (int l = 0; l < NC; l++) {
            count[l] = 0;
            for (int r = 0; r < 3; r++) { centers_b[l][r] = 0.f; fix_b[l][r] = 0; }
            for (int li = 0; li < NC; li++) {
                cluster_distances[l][li].idx = li;
                cluster_distances[l][li].distance = sqnorm3(centers_a[l], centers_a[li]);
            }
            sortDistances(cluster_distances[l]);
        }
        for (int u = 0; u < cols_km; u++)
            for (int v = 0; v < rows_km; v++)
                if (depth_ref(v, u) != 0.f) {
                    const int last_label = labels_lowres(v, u);
                    int best_label = last_label;
                    const IndexAndDistance* distances = cluster_distances[last_label];
                    const float pnt[3] = {depth_ref(v, u), xx_ref(v, u), yy_ref(v, u)};
                    const float distance_to_last_label = sqnorm3(centers_a[last_label], pnt);
                    float best_distance = distance_to_last_label;
                    for (int li = 1; li < NC; ++li) {
                        if (distances[li].distance > 4.f * distance_to_last_label) break;
                        const float distance_to_label = sqnorm3(centers_a[distances[li].idx], pnt);
                        if (distance_to_label < best_distance) { best_distance = distance_to_label; best_label = distances[li].idx; }
                    }
                    labels_lowres(v, u) = best_label;
                    if (accum == ORC_ACCUM_F32)
                        for (int r = 0; r < 3; r++) centers_b[best_label][r] += pnt[r];
                    else
                        for (int r = 0; r < 3; r++) fix_b[best_label][r] += fixq(pnt[r], 36);
                    count[best_label] += 1;
                }
        for (int l = 0; l < NC; l++)
            if (count[l] > 0)
                for (int r = 0; r < 3; r++) {
                    if (accum == ORC_ACCUM_F32) centers_b[l][r] /= float(count[l]);
                    else centers_b[l][r] = (float)(fixval(fix_b[l][r], 36) / (double)count[l]);
                }
        float max_diff = 0.f; /* :224 */
        for (int l = 0; l < NC; l++)
            for (int r = 0; r < 3; r++) max_diff = std::max(max_diff, std::fabs(centers_a[l][r] - centers_b[l][r]));
        std::memcpy(centers_a, centers_b, sizeof(centers_a));
        if (max_diff < 1e-2f) break;
    }
    for (int c = 0; c < NC; c++)
        for (int r = 0; r < 3; r++) kmeans[r][c] = centers_a[c][r];

    /* :238-291 full resolution labelling */
    const Img& depth_highres = depthPyr[0];
    const Img& xx_highres = xxPyr[0];
    const Img& yy_highres = yyPyr[0];
    ImgI& labels_ref = clusterAllocation[0];
    labels_ref.fill(NC);
    for (int l = 0; l < NC; l++) {
        for (int li = 0; li < NC; li++) {
            cluster_distances[l][li].idx = li;
            cluster_distances[l][li].distance = sqnorm3(centers_a[l], centers_a[li]);
        }
        sortDistances(cluster_distances[l]);
    }
    for (int u = 0; u < cols; u++)
        for (int v = 0; v < rows; v++)
            if (depth_highres(v, u) != 0.f) {
                const int label_lowres_here = labels_lowres(v / 2, u / 2);
                const int last_label = (label_lowres_here == NC) ? 0 : label_lowres_here;
                int best_label = last_label;
                const IndexAndDistance* distances = cluster_distances[last_label];
                const float pnt[3] = {depth_highres(v, u), xx_highres(v, u), yy_highres(v, u)};
                const float distance_to_last_label = sqnorm3(centers_a[last_label], pnt);
                float best_distance = distance_to_last_label;
                for (int li = 1; li < NC; ++li) {
                    if (distances[li].distance > 4.f * distance_to_last_label) break;
                    const float distance_to_label = sqnorm3(centers_a[distances[li].idx], pnt);
                    if (distance_to_label < best_distance) { best_distance = distance_to_label; best_label = distances[li].idx; }
                }
                labels_ref(v, u) = best_label;
            }
    computeRegionConnectivity();
}

/* KMeans.cpp:297-341 */
void orc_ctx::computeRegionConnectivity() {
    const float dist2_threshold = sq(0.03f * 120.f / float(rows));
    const ImgI& labels_ref = clusterAllocation[0];
    const Img& depth_ref = depthPyr[0];
    const Img& xx_ref = xxPyr[0];
    const Img& yy_ref = yyPyr[0];
    for (int i = 0; i < NC; i++)
        for (int j = 0; j < NC; j++) connectivity[i][j] = (i == j);
    for (int u = 0; u < cols - 1; u++)
        for (int v = 0; v < rows - 1; v++)
            if (depth_ref(v, u) != 0.f) {
                if ((labels_ref(v, u) != labels_ref(v + 1, u)) && (labels_ref(v + 1, u) != NC)) {
                    const float disty = sq(depth_ref(v, u) - depth_ref(v + 1, u)) + sq(yy_ref(v, u) - yy_ref(v + 1, u));
                    if (disty < dist2_threshold) {
                        connectivity[labels_ref(v, u)][labels_ref(v + 1, u)] = true;
                        connectivity[labels_ref(v + 1, u)][labels_ref(v, u)] = true;
                    }
                }
                if ((labels_ref(v, u) != labels_ref(v, u + 1)) && (labels_ref(v, u + 1) != NC)) {
                    const float distx = sq(depth_ref(v, u) - depth_ref(v, u + 1)) + sq(xx_ref(v, u) - xx_ref(v, u + 1));
                    if (distx < dist2_threshold) {
                        connectivity[labels_ref(v, u)][labels_ref(v, u + 1)] = true;
                        connectivity[labels_ref(v, u + 1)][labels_ref(v, u)] = true;
                    }
                }
            }
}

/* KMeans.cpp:343-391 */
void orc_ctx::createClustersPyramidUsingKMeans() {
    float kmeans_dist[NC][NC];
    float cen[NC][3];
    for (int l = 0; l < NC; l++)
        for (int r = 0; r < 3; r++) cen[l][r] = kmeans[r][l];
    for (int la = 0; la < NC; la++)
        for (int lb = la + 1; lb < NC; lb++) kmeans_dist[la][lb] = sqnorm3(cen[la], cen[lb]);
    for (int i = 2; i < p.ctf_levels; i++) {
        const int cols_km = cols >> i, rows_km = rows >> i;
        ImgI& labels_ref = clusterAllocation[i];
        const Img& depth_old_ref = depthPyr[i];
        const Img& xx_old_ref = xxPyr[i];
        const Img& yy_old_ref = yyPyr[i];
        labels_ref.fill(NC);
        for (int u = 0; u < cols_km; u++)
            for (int v = 0; v < rows_km; v++)
                if (depth_old_ref(v, u) != 0.f) {
                    int label = 0;
                    const float pnt[3] = {depth_old_ref(v, u), xx_old_ref(v, u), yy_old_ref(v, u)};
                    float min_dist = sqnorm3(cen[0], pnt);
                    float dist_here;
                    for (int l = 1; l < NC; l++) {
                        if (kmeans_dist[label][l] > 4.f * min_dist) continue;
                        else if ((dist_here = sqnorm3(cen[l], pnt)) < min_dist) { label = l; min_dist = dist_here; }
                    }
                    labels_ref(v, u) = label;
                }
    }
}

/* FrontEnd.cpp:775-892 */
void orc_ctx::warpImagesAccurateInverse() {
    const float f = float(cols_i) / (2.f * std::tan(0.5f * p.fovh));
    const float disp_u_i = 0.5f * float(cols_i - 1);
    const float disp_v_i = 0.5f * float(rows_i - 1);
    Img& depth_warped_ref = depthWarpedPyr[image_level];
    Img& intensity_warped_ref = intensityWarpedPyr[image_level];
    Img& xx_warped_ref = xxWarpedPyr[image_level];
    Img& yy_warped_ref = yyWarpedPyr[image_level];
    const Img& depth_ref = depthPredPyr[image_level];
    const Img& intensity_ref = intensityPredPyr[image_level];
    const Img& xx_ref = xxPredPyr[image_level];
    const Img& yy_ref = yyPredPyr[image_level];
    depth_warped_ref.fill(0.f);
    intensity_warped_ref.fill(0.f);
    const size_t np = (size_t)rows_i * cols_i;
    std::vector<float> wacu(np, 0.f);
    std::vector<int64_t> dfix, ifix;
    std::vector<int32_t> wfix;
    const bool exact = (accum != ORC_ACCUM_F32);
    if (exact) { dfix.assign(np, 0); ifix.assign(np, 0); wfix.assign(np, 0); }
    const int cols_lim = 100 * (cols_i - 1);
    const int rows_lim = 100 * (rows_i - 1);
    float T[16];
    if (exact) rigid_inverse<double>(T_odometry, T); else general_inverse<float>(4, T_odometry, T); /* :800 */

    auto splat = [&](int v, int u, int w, float depth_w, float intensity_w) {
        const size_t k = (size_t)v * cols_i + u;
        if (exact) {
            dfix[k] += (int64_t)w * fixq(depth_w, 32);
            ifix[k] += (int64_t)w * fixq(intensity_w, 22);
            wfix[k] += w;
        } else {
            depth_warped_ref.a[k] += float(w) * depth_w;
            intensity_warped_ref.a[k] += float(w) * intensity_w;
            wacu[k] += float(w);
        }
    };

    for (int j = 0; j < cols_i; j++)
        for (int i = 0; i < rows_i; i++) {
            const float z = depth_ref(i, j);
            if (z != 0.f) {
                const float intensity_w = intensity_ref(i, j);
                const float xr = xx_ref(i, j), yr = yy_ref(i, j);
                const float x_w = T[0] * xr + T[1] * yr + T[2] * z + T[3];
                const float y_w = T[4] * xr + T[5] * yr + T[6] * z + T[7];
                const float depth_w = T[8] * xr + T[9] * yr + T[10] * z + T[11];
                const float fu = 100.f * (f * x_w / depth_w + disp_u_i);
                const float fv = 100.f * (f * y_w / depth_w + disp_v_i);
                /* int() of a non-finite or out-of-range float is UB in C++ (x86 yields INT_MIN, failing
                 * the >= 0 test below); restated explicitly as "out of bounds" */
                if (!(std::fabs(fu) < 1.0e9f) || !(std::fabs(fv) < 1.0e9f)) continue;
                const int uwarp = int(fu);
                const int vwarp = int(fv);
                if ((uwarp >= 0) && (uwarp < cols_lim) && (vwarp >= 0) && (vwarp < rows_lim)) {
                    const int uwarp_l = uwarp - uwarp % 100;
                    const int uwarp_r = uwarp_l + 100;
                    const int vwarp_d = vwarp - vwarp % 100;
                    const int vwarp_u = vwarp_d + 100;
                    const int delta_r = uwarp_r - uwarp;
                    const int delta_l = 100 - delta_r;
                    const int delta_u = vwarp_u - vwarp;
                    const int delta_d = 100 - delta_u;
                    if (std::min(delta_r, delta_l) + std::min(delta_u, delta_d) < 5) {
                        const int ind_u = delta_r > delta_l ? uwarp_l / 100 : uwarp_r / 100;
                        const int ind_v = delta_u > delta_d ? vwarp_d / 100 : vwarp_u / 100;
                        splat(ind_v, ind_u, 200, depth_w, intensity_w);
                    } else {
                        const int v_d = vwarp_d / 100, u_l = uwarp_l / 100;
                        const int v_u = v_d + 1, u_r = u_l + 1;
                        splat(v_u, u_r, delta_l + delta_d, depth_w, intensity_w);
                        splat(v_u, u_l, delta_r + delta_d, depth_w, intensity_w);
                        splat(v_d, u_r, delta_l + delta_u, depth_w, intensity_w);
                        splat(v_d, u_l, delta_r + delta_u, depth_w, intensity_w);
                    }
                }
            }
        }

    const float inv_f_i = 1.f / f;
    for (int u = 0; u < cols_i; u++)
        for (int v = 0; v < rows_i; v++) {
            const size_t k = (size_t)v * cols_i + u;
            bool hit;
            if (exact) {
                hit = wfix[k] != 0;
                if (hit) {
                    intensity_warped_ref.a[k] = (float)((double)ifix[k] / ((double)wfix[k] * 4194304.0));
                    depth_warped_ref.a[k] = (float)((double)dfix[k] / ((double)wfix[k] * 4294967296.0));
                }
            } else {
                hit = wacu[k] != 0.f;
                if (hit) {
                    intensity_warped_ref.a[k] /= wacu[k];
                    depth_warped_ref.a[k] /= wacu[k];
                }
            }
            if (hit) {
                xx_warped_ref(v, u) = (float(u) - disp_u_i) * depth_warped_ref(v, u) * inv_f_i;
                yy_warped_ref(v, u) = (float(v) - disp_v_i) * depth_warped_ref(v, u) * inv_f_i;
            } else {
                xx_warped_ref(v, u) = 0.f;
                yy_warped_ref(v, u) = 0.f;
            }
        }
}

/* FrontEnd.cpp:393-430 */
void orc_ctx::calculateCoord() {
    validPixels.clear();
    validPixels.reserve((size_t)rows_i * cols_i);
    std::fill(Null.begin(), Null.end(), 0);
    Img& depth_inter_ref = depthInterPyr[image_level];
    Img& xx_inter_ref = xxInterPyr[image_level];
    Img& yy_inter_ref = yyInterPyr[image_level];
    Img& intensity_inter_ref = intensityInterPyr[image_level];
    const Img& depth_ref = depthPyr[image_level];
    const Img& depth_warped_ref = depthWarpedPyr[image_level];
    for (int u = 0; u != cols_i; u++)
        for (int v = 0; v != rows_i; v++) {
            if ((depth_ref(v, u) != 0.f) && (depth_warped_ref(v, u) != 0.f)) {
                depth_inter_ref(v, u) = 0.5f * (depth_ref(v, u) + depth_warped_ref(v, u));
                xx_inter_ref(v, u) = 0.5f * (xxPyr[image_level](v, u) + xxWarpedPyr[image_level](v, u));
                yy_inter_ref(v, u) = 0.5f * (yyPyr[image_level](v, u) + yyWarpedPyr[image_level](v, u));
                if ((u != 0) && (v != 0) && (u != cols_i - 1) && (v != rows_i - 1)) validPixels.push_back(std::make_pair(v, u));
            } else {
                Null[(size_t)v * cols_i + u] = 1;
                depth_inter_ref(v, u) = 0.f;
                xx_inter_ref(v, u) = 0.f;
                yy_inter_ref(v, u) = 0.f;
            }
            intensity_inter_ref(v, u) = 0.5f * (intensityPyr[image_level](v, u) + intensityWarpedPyr[image_level](v, u));
        }
}

/* FrontEnd.cpp:432-479.  dcu..ddv/dct/ddt are kept full-res with the active level in the top-left block. */
void orc_ctx::calculateDerivatives() {
    Img rx, ry, rx_intensity, ry_intensity;
    rx.resize(rows_i, cols_i); ry.resize(rows_i, cols_i);
    rx_intensity.resize(rows_i, cols_i); ry_intensity.resize(rows_i, cols_i);
    rx.fill(1.f); ry.fill(1.f); rx_intensity.fill(1.f); ry_intensity.fill(1.f);
    const Img& depth_ref = depthInterPyr[image_level];
    const Img& intensity_ref = intensityInterPyr[image_level];
    const float epsilon_intensity = 1e-6f;
    const float epsilon_depth = 0.005f;
    auto isNull = [&](int v, int u) { return Null[(size_t)v * cols_i + u] != 0; };
    for (int u = 0; u < cols_i - 1; u++)
        for (int v = 0; v < rows_i; v++)
            if (!isNull(v, u)) {
                rx(v, u) = std::fabs(depth_ref(v, u + 1) - depth_ref(v, u)) + epsilon_depth;
                rx_intensity(v, u) = std::fabs(intensity_ref(v, u + 1) - intensity_ref(v, u)) + epsilon_intensity;
            }
    for (int u = 0; u < cols_i; u++)
        for (int v = 0; v < rows_i - 1; v++)
            if (!isNull(v, u)) {
                ry(v, u) = std::fabs(depth_ref(v + 1, u) - depth_ref(v, u)) + epsilon_depth;
                ry_intensity(v, u) = std::fabs(intensity_ref(v + 1, u) - intensity_ref(v, u)) + epsilon_intensity;
            }
    for (int v = 1; v < rows_i - 1; v++)
        for (int u = 1; u < cols_i - 1; u++)
            if (!isNull(v, u)) {
                dcu(v, u) = (rx_intensity(v, u - 1) * (intensity_ref(v, u + 1) - intensity_ref(v, u)) + rx_intensity(v, u) * (intensity_ref(v, u) - intensity_ref(v, u - 1))) / (rx_intensity(v, u) + rx_intensity(v, u - 1));
                ddu(v, u) = (rx(v, u - 1) * (depth_ref(v, u + 1) - depth_ref(v, u)) + rx(v, u) * (depth_ref(v, u) - depth_ref(v, u - 1))) / (rx(v, u) + rx(v, u - 1));
                dcv(v, u) = (ry_intensity(v - 1, u) * (intensity_ref(v + 1, u) - intensity_ref(v, u)) + ry_intensity(v, u) * (intensity_ref(v, u) - intensity_ref(v - 1, u))) / (ry_intensity(v, u) + ry_intensity(v - 1, u));
                ddv(v, u) = (ry(v - 1, u) * (depth_ref(v + 1, u) - depth_ref(v, u)) + ry(v, u) * (depth_ref(v, u) - depth_ref(v - 1, u))) / (ry(v, u) + ry(v - 1, u));
            }
    for (int v = 0; v < rows_i; v++)
        for (int u = 0; u < cols_i; u++) {
            dct(v, u) = intensityPyr[image_level](v, u) - intensityWarpedPyr[image_level](v, u);
            ddt(v, u) = depthPyr[image_level](v, u) - depthWarpedPyr[image_level](v, u);
        }
}

/* FrontEnd.cpp:481-510 */
void orc_ctx::computeWeights() {
    weights_c.fill(0.f);
    weights_d.fill(0.f);
    const float kduvt_c = 10.f, kduvt_d = 200.f;
    const float error_m_c = 1.f, error_m_d = 0.01f;
    for (auto i : validPixels) {
        const int v = i.first, u = i.second;
        const float error_l_c = kduvt_c * (std::fabs(dct(v, u)) + std::fabs(dcu(v, u)) + std::fabs(dcv(v, u)));
        const float error_l_d = kduvt_d * (std::fabs(ddt(v, u)) + std::fabs(ddu(v, u)) + std::fabs(ddv(v, u)));
        weights_c(v, u) = sqrtf(1.f / (error_m_c + error_l_c));
        weights_d(v, u) = sqrtf(1.f / (error_m_d + error_l_d));
    }
    float mc = 0.f, md = 0.f;
    for (float x : weights_c.a) mc = std::max(mc, x);
    for (float x : weights_d.a) md = std::max(md, x);
    max_wc_raw = mc; max_wd_raw = md;
    /* the raw weights stay in weights_c/d; the 1/max normalisation (:505-509) is applied where they are read */
}

/* SegmentationBackground.cpp:53-103 */
void orc_ctx::computeSegPrior() {
    int cluster_size[NC], cluster_nonnull[NC];
    int64_t fix_prior[NC];
    const ImgI& labels_ref = clusterAllocation[image_level];
    for (int l = 0; l < NC; l++) { b_prior[l] = 0.f; cluster_size[l] = 0; cluster_nonnull[l] = 0; lambda_t_w[l] = 0.f; fix_prior[l] = 0; }
    for (int u = 0; u < cols_i; u++)
        for (int v = 0; v < rows_i; v++) {
            const int l = labels_ref(v, u);
            if (l != NC) {
                if (!Null[(size_t)v * cols_i + u]) {
                    cluster_nonnull[l]++;
                    const float term = 1.f - p.kz * std::fabs(ddt(v, u));
                    if (accum == ORC_ACCUM_F32) b_prior[l] += term;
                    else fix_prior[l] += fixq(term, 32);
                }
                cluster_size[l]++;
            }
        }
    for (int l = 0; l < NC; l++)
        if (cluster_size[l] != 0) {
            const float ratio = float(cluster_nonnull[l]) / float(cluster_size[l]);
            if (ratio < 0.1f) {
                lambda_t_w[l] = 0.1f;
                b_prior[l] = -1.f;
            } else {
                lambda_t_w[l] = ratio;
                const float mean = (accum == ORC_ACCUM_F32) ? b_prior[l] / float(cluster_nonnull[l])
                                                            : (float)(fixval(fix_prior[l], 32) / (double)cluster_nonnull[l]);
                b_prior[l] = std::max(-1.f, std::min(2.f, mean));
            }
        }
}

/* SegmentationBackground.cpp:105-174.  A_seg = [diag(a_l); +-2*lambda_reg rows] so
 * AtA_seg = diag(a_l^2) + 4*lambda_reg^2 * Laplacian, AtB_seg = a_l * B_l. */
void orc_ctx::solveSegmIteration(const float* aver_res, float aver_res_overall, const double* lap) {
    const float kc = p.kc_cauchy;
    if (accum == ORC_ACCUM_F32) {
        /* literal: A_seg is (24 + nc) x 24 with one (+2 lambda_reg, -2 lambda_reg) row per adjacent pair l < lc
         * (SegmentationBackground.cpp:105-130); AtA_seg / AtB_seg are formed by sequential sums over its rows (:164-165) */
        std::vector<std::pair<int, int>> conn;
        for (int l = 0; l < NC; l++)
            for (int lc = l + 1; lc < NC; lc++)
                if (lap[l * NC + lc] != 0.0) conn.push_back(std::make_pair(l, lc));
        const int nr = NC + (int)conn.size();
        std::vector<float> As((size_t)nr * NC, 0.f), Bs(nr, 0.f);
        const float weight_reg = 2.f * p.lambda_reg;
        for (size_t q = 0; q < conn.size(); q++) { As[(NC + q) * NC + conn[q].first] = weight_reg; As[(NC + q) * NC + conn[q].second] = -weight_reg; }
        const float repr_res = std::max(0.001f, aver_res_overall);
        const float fixed_term = std::log(1.f + sq(p.kb * repr_res / (kc * aver_res_overall)));
        const float mult_res = 1.f / (kc * aver_res_overall);
        for (int l = 0; l < NC; l++) {
            if (lambda_t_w[l] > 0.1f) {
                const float dataterm = fixed_term - std::log(1.f + sq(aver_res[l] * mult_res));
                As[(size_t)l * NC + l] = 2.f * lambda_t_w[l] * p.lambda_prior;
                Bs[l] = dataterm + 2.f * p.lambda_prior * lambda_t_w[l] * b_prior[l];
            } else {
                As[(size_t)l * NC + l] = 2.f * lambda_t_w[l];
                Bs[l] = 2.f * lambda_t_w[l] * b_prior[l];
            }
        }
        float A[NC * NC], rhs[NC], x[NC];
        unsigned char zero[NC];
        for (int i = 0; i < NC; i++) {
            for (int j = 0; j < NC; j++) { float acc = 0.f; for (int r = 0; r < nr; r++) acc += As[(size_t)r * NC + i] * As[(size_t)r * NC + j]; A[i * NC + j] = acc; }
            float acc = 0.f;
            for (int r = 0; r < nr; r++) acc += As[(size_t)r * NC + i] * Bs[r];
            rhs[i] = acc;
        }
        ldlt_factor<float>(NC, A, zero);
        ldlt_solve_factored<float>(NC, A, zero, rhs, x);
        for (int l = 0; l < NC; l++) b_segm[l] = std::max(-1.f, std::min(2.f, x[l]));
    } else {
        double A[NC * NC], rhs[NC], x[NC];
        unsigned char zero[NC];
        const double aver = (double)aver_res_overall;
        const double repr_res = (double)std::max(0.001f, aver_res_overall);
        const double r0 = (double)p.kb * repr_res / ((double)kc * aver);
        const double fixed_term = std::log(1.0 + r0 * r0);
        const double mult_res = 1.0 / ((double)kc * aver);
        const double wreg = 2.0 * (double)p.lambda_reg;
        const double wreg2 = wreg * wreg;
        for (int i = 0; i < NC * NC; i++) A[i] = wreg2 * lap[i];
        for (int l = 0; l < NC; l++) {
            double a, b;
            const double ltw = (double)lambda_t_w[l];
            if (lambda_t_w[l] > 0.1f) {
                const double rl = (double)aver_res[l] * mult_res;
                const double dataterm = fixed_term - std::log(1.0 + rl * rl);
                a = 2.0 * ltw * (double)p.lambda_prior;
                b = dataterm + 2.0 * (double)p.lambda_prior * ltw * (double)b_prior[l];
            } else {
                a = 2.0 * ltw;
                b = 2.0 * ltw * (double)b_prior[l];
            }
            A[l * NC + l] += a * a;
            rhs[l] = a * b;
        }
        ldlt_factor<double>(NC, A, zero);
        ldlt_solve_factored<double>(NC, A, zero, rhs, x);
        for (int l = 0; l < NC; l++) b_segm[l] = (float)std::max(-1.0, std::min(2.0, x[l]));
    }
}

/* FrontEnd.cpp:513-692 */
void orc_ctx::solveOdometryAndSegmJoint() {
    /* buildSystemSegm (SegmentationBackground.cpp:105-130): graph Laplacian of the l<lc adjacency */
    double lap[NC * NC];
    for (int i = 0; i < NC * NC; i++) lap[i] = 0;
    if (p.enable_segmentation)
        for (int l = 0; l < NC; l++)
            for (int lc = l + 1; lc < NC; lc++)
                if (connectivity[l][lc]) {
                    lap[l * NC + l] += 1; lap[lc * NC + lc] += 1;
                    lap[l * NC + lc] -= 1; lap[lc * NC + l] -= 1;
                }

    const ImgI& labels_ref = clusterAllocation[image_level];
    const Img& depth_inter_ref = depthInterPyr[image_level];
    const Img& xx_inter_ref = xxInterPyr[image_level];
    const Img& yy_inter_ref = yyInterPyr[image_level];
    const size_t N = validPixels.size();
    const bool exact = (accum != ORC_ACCUM_F32);   /* fixed-point bounded sums + double algebra */
    const bool quant = (accum == ORC_ACCUM_EXACT); /* integer normal equations and |res|^2 (the CUDA contract) */
    float* tr = cur_trace;
    tr[3] = (float)N;
    tr[5] = max_wc_raw; tr[6] = max_wd_raw;

    if (N == 0 || !(max_wc_raw > 0.f) || !(max_wd_raw > 0.f)) { /* undefined in the reference */
        status |= 1;
        for (int i = 0; i < 6; i++) twist_level_odometry[i] = 0.f;
        return;
    }
    const float inv_max_c = 1.f / max_wc_raw, inv_max_d = 1.f / max_wd_raw;

    std::vector<float> A(2 * N * 6), B(2 * N), res(2 * N);
    /* EXACT / F64 policies: rows built with the RAW pre-weights; the 1/max normalisation (:505-509) is applied to the
     * residual and folded into the robust weight instead of into every row entry (one rounding different per entry,
     * far below the integer quantisation; lets the CUDA path store the rows once per linearisation). */
    std::vector<float> Araw, Braw;
    std::vector<int> lab(N);
    const float f_inv = float(cols_i) / (2.f * std::tan(0.5f * p.fovh));
    /* the two Jacobian rows of a pixel, :545-585; wc_n / wd_n are the (normalised) pre-weights */
    auto build_rows = [&](int v, int u, float wc_n, float wd_n, float* ac, float& bc, float* ad, float& bd) {
        const float d = depth_inter_ref(v, u);
        const float inv_d = 1.f / d;
        const float x = xx_inter_ref(v, u);
        const float y = yy_inter_ref(v, u);
        const float dycomp_c = dcu(v, u) * f_inv * inv_d;
        const float dzcomp_c = dcv(v, u) * f_inv * inv_d;
        const float twc = wc_n * p.k_photometric_res;
        ac[0] = twc * (-dycomp_c);
        ac[1] = twc * (-dzcomp_c);
        ac[2] = twc * (dycomp_c * x * inv_d + dzcomp_c * y * inv_d);
        ac[3] = twc * (dycomp_c * inv_d * y * x + dzcomp_c * (y * y * inv_d + d));
        ac[4] = twc * (-dycomp_c * (x * x * inv_d + d) - dzcomp_c * inv_d * y * x);
        ac[5] = twc * (dycomp_c * y - dzcomp_c * x);
        bc = twc * (-dct(v, u));
        const float dycomp_d = ddu(v, u) * f_inv * inv_d;
        const float dzcomp_d = ddv(v, u) * f_inv * inv_d;
        const float twd = wd_n;
        ad[0] = twd * (-dycomp_d);
        ad[1] = twd * (-dzcomp_d);
        ad[2] = twd * (1.f + dycomp_d * x * inv_d + dzcomp_d * y * inv_d);
        ad[3] = twd * (y + dycomp_d * inv_d * y * x + dzcomp_d * (y * y * inv_d + d));
        ad[4] = twd * (-x - dycomp_d * (x * x * inv_d + d) - dzcomp_d * inv_d * y * x);
        ad[5] = twd * (dycomp_d * y - dzcomp_d * x);
        bd = twd * (-ddt(v, u));
    };
    size_t cont = 0;
    int64_t fixBc = 0, fixBd = 0;
    float Mc[7] = {0, 0, 0, 0, 0, 0, 0}, Md[7] = {0, 0, 0, 0, 0, 0, 0}; /* column maxima of the rows built with raw weights */
    for (size_t n = 0; n < N; n++) {
        const int v = validPixels[n].first, u = validPixels[n].second;
        lab[n] = p.enable_segmentation ? labels_ref(v, u) : 0;
        const float wc_n = inv_max_c * weights_c(v, u); /* :505-509 */
        const float wd_n = inv_max_d * weights_d(v, u);
        build_rows(v, u, wc_n, wd_n, &A[cont * 6], B[cont], &A[(cont + 1) * 6], B[cont + 1]);
        cont += 2;
        if (exact) {
            fixBc += fixq(weights_c(v, u) * std::fabs(dct(v, u)), 32);
            fixBd += fixq(weights_d(v, u) * std::fabs(ddt(v, u)), 32);
            if (Araw.empty()) { Araw.resize(2 * N * 6); Braw.resize(2 * N); }
            float* ac = &Araw[(cont - 2) * 6];
            float* ad = &Araw[(cont - 1) * 6];
            float& bc = Braw[cont - 2];
            float& bd = Braw[cont - 1];
            build_rows(v, u, weights_c(v, u), weights_d(v, u), ac, bc, ad, bd);
            for (int c = 0; c < 6; c++) { Mc[c] = std::max(Mc[c], std::fabs(ac[c])); Md[c] = std::max(Md[c], std::fabs(ad[c])); }
            Mc[6] = std::max(Mc[6], std::fabs(bc)); Md[6] = std::max(Md[6], std::fabs(bd));
        }
    }
    /* power-of-two column scales of the EXACT policy */
    int sexp[7] = {0, 0, 0, 0, 0, 0, 0};
    float colbound[7] = {0, 0, 0, 0, 0, 0, 0};
    if (exact)
        for (int c = 0; c < 7; c++) {
            colbound[c] = std::max(inv_max_c * Mc[c], inv_max_d * Md[c]);
            sexp[c] = scale_exponent(colbound[c]);
        }
    /* :589-590 */
    float aver_res;
    if (!exact) {
        float s = 0.f;
        for (size_t i = 0; i < 2 * N; i++) { res[i] = -B[i]; s += std::fabs(res[i]); }
        aver_res = s / float(2 * N);
    } else {
        for (size_t i = 0; i < 2 * N; i++) res[i] = ((i & 1) ? inv_max_d : inv_max_c) * (-Braw[i]);
        const double sc = (double)inv_max_c * (double)p.k_photometric_res;
        aver_res = (float)((sc * fixval(fixBc, 32) + (double)inv_max_d * fixval(fixBd, 32)) / (double)(2 * N));
    }
    tr[7] = aver_res;
    if (!(aver_res > 0.f) || !std::isfinite(aver_res)) { /* identical images: 1/(kc*0) -> NaN in the reference */
        status |= 2;
        for (int i = 0; i < 6; i++) twist_level_odometry[i] = 0.f;
        return;
    }

    float Var[6] = {0, 0, 0, 0, 0, 0}, prev_sol[6] = {0, 0, 0, 0, 0, 0};
    if (!p.enable_segmentation) for (int l = 0; l < NC; l++) b_segm[l] = 1.f; /* :607 variant */
    else if (level == 0) for (int l = 0; l < NC; l++) b_segm[l] = b_prior[l];  /* :603-604 */

    double AtA[36], AtB[6];
    double res_sq = 0;
    int it_done = 0;
    for (int k = 1; k <= p.max_iter_irls; k++) {
        const float inv_c_Cauchy = 1.f / (p.kc_cauchy * aver_res);
        for (int i = 0; i < 36; i++) AtA[i] = 0;
        for (int i = 0; i < 6; i++) AtB[i] = 0;
        float AtAf[36], AtBf[6];
        int64_t AtAq[36], AtBq[6];
        for (int i = 0; i < 36; i++) { AtAf[i] = 0.f; AtAq[i] = 0; }
        for (int i = 0; i < 6; i++) { AtBf[i] = 0.f; AtBq[i] = 0; }
        for (size_t n = 0; n < N; n++) {
            const float b_weight = std::max(0.f, std::min(1.f, b_segm[lab[n]]));
            for (int r = 0; r < 2; r++) { /* :627-636 */
                const size_t row = 2 * n + r;
                const float w = b_weight * sqrtf(1.f / (1.f + sq(res[row] * inv_c_Cauchy)));
                float aw[6], bw;
                if (exact) { /* normalisation folded into the weight: (w * inv_max) * raw entry */
                    const float wm = w * (r ? inv_max_d : inv_max_c);
                    for (int c = 0; c < 6; c++) aw[c] = wm * Araw[row * 6 + c];
                    bw = wm * Braw[row];
                } else {
                    for (int c = 0; c < 6; c++) aw[c] = w * A[row * 6 + c];
                    bw = w * B[row];
                }
                if (quant) {
                    /* (w * (inv_max * 2^s)) * raw entry: the power of two commutes with the rounding */
                    float aws[6];
                    const float im = r ? inv_max_d : inv_max_c;
                    for (int c = 0; c < 6; c++) aws[c] = (w * ldexpf(im, sexp[c])) * Araw[row * 6 + c];
                    const float bws = (w * ldexpf(im, sexp[6])) * Braw[row];
                    for (int i = 0; i < 6; i++) {
                        for (int j = i; j < 6; j++) {
                            const int64_t q = qprod(aws[i], aws[j]);
                            if (q > ((int64_t)1 << (2 * ORC_QSCALE_BITS + ORC_QFRAC_BITS)) || q < -((int64_t)1 << (2 * ORC_QSCALE_BITS + ORC_QFRAC_BITS))) status |= 8; /* scale bound violated: must never happen */
                            AtAq[i * 6 + j] += q;
                        }
                        AtBq[i] += qprod(aws[i], bws);
                    }
                } else if (exact) {
                    for (int i = 0; i < 6; i++) {
                        for (int j = i; j < 6; j++) AtA[i * 6 + j] += (double)aw[i] * (double)aw[j];
                        AtB[i] += (double)aw[i] * (double)bw;
                    }
                } else {
                    for (int i = 0; i < 6; i++) {
                        for (int j = i; j < 6; j++) AtAf[i * 6 + j] += aw[i] * aw[j];
                        AtBf[i] += aw[i] * bw;
                    }
                }
            }
        }
        if (!exact) {
            for (int i = 0; i < 36; i++) AtA[i] = AtAf[i];
            for (int i = 0; i < 6; i++) AtB[i] = AtBf[i];
        }
        if (quant) {
            for (int i = 0; i < 6; i++) {
                for (int j = i; j < 6; j++) AtA[i * 6 + j] = std::ldexp((double)AtAq[i * 6 + j], -(sexp[i] + sexp[j]) - ORC_QFRAC_BITS);
                AtB[i] = std::ldexp((double)AtBq[i], -(sexp[i] + sexp[6]) - ORC_QFRAC_BITS);
            }
        }
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < i; j++) AtA[i * 6 + j] = AtA[j * 6 + i];
        /* :642 Var = AtA.ldlt().solve(AtB) */
        int nz;
        if (exact) {
            double F[36], x[6]; unsigned char zero[6];
            for (int i = 0; i < 36; i++) F[i] = AtA[i];
            nz = ldlt_factor<double>(6, F, zero);
            ldlt_solve_factored<double>(6, F, zero, AtB, x);
            for (int i = 0; i < 6; i++) Var[i] = (float)x[i];
        } else {
            float F[36], bb[6], x[6]; unsigned char zero[6];
            for (int i = 0; i < 36; i++) F[i] = (float)AtA[i];
            for (int i = 0; i < 6; i++) bb[i] = (float)AtB[i];
            nz = ldlt_factor<float>(6, F, zero);
            ldlt_solve_factored<float>(6, F, zero, bb, x);
            for (int i = 0; i < 6; i++) Var[i] = x[i];
        }
        if (nz) status |= 4;
        /* :644-646 */
        for (size_t i = 0; i < 2 * N; i++) {
            if (exact) {
                float r = -Braw[i];
                for (int c = 0; c < 6; c++) r += Var[c] * Araw[i * 6 + c];
                res[i] = ((i & 1) ? inv_max_d : inv_max_c) * r;
            } else {
                float r = -B[i];
                for (int c = 0; c < 6; c++) r += Var[c] * A[i * 6 + c];
                res[i] = r;
            }
        }
        /* :650-667 */
        float aver_res_label[NC];
        int num_pix_label[NC];
        int64_t fix_label[NC];
        for (int l = 0; l < NC; l++) { aver_res_label[l] = 0.f; num_pix_label[l] = 1; fix_label[l] = 0; }
        const float aver_res_old = aver_res;
        double rs_d = 0; float rs_f = 0.f;
        /* EXACT: |res| <= |B| + sum_k |Var_k| |A_k| bounds the residuals; same power-of-two scaling as above */
        int rexp = 0, lexp = 0; int64_t rs_q = 0;
        if (quant) {
            float rb = colbound[6];
            for (int c = 0; c < 6; c++) rb += std::fabs(Var[c]) * colbound[c];
            rexp = scale_exponent(rb);
            lexp = rexp + (ORC_LABEL_BITS - ORC_QSCALE_BITS - 1); /* |res_c| + |res_d| < 2^(11 - rexp): the scaled term stays below 2^30 */
        }
        for (size_t n = 0; n < N; n++) {
            const float ress_here = std::fabs(res[2 * n]) + std::fabs(res[2 * n + 1]);
            if (quant) fix_label[lab[n]] += (int64_t)llrintf(ldexpf(ress_here, lexp)); /* cvt.rni.s32.f32 of the scaled term on the GPU */
            else if (exact) fix_label[lab[n]] += fixq(ress_here, 30);
            else aver_res_label[lab[n]] += ress_here;
            num_pix_label[lab[n]]++;
            if (quant) {
                const float r0 = ldexpf(res[2 * n], rexp), r1 = ldexpf(res[2 * n + 1], rexp);
                rs_q += qprod(r0, r0) + qprod(r1, r1);
            } else if (exact) rs_d += (double)res[2 * n] * (double)res[2 * n] + (double)res[2 * n + 1] * (double)res[2 * n + 1];
            else { rs_f += res[2 * n] * res[2 * n]; rs_f += res[2 * n + 1] * res[2 * n + 1]; }
        }
        if (exact) {
            int64_t tot = 0;
            const int ls = quant ? lexp : 30;
            for (int l = 0; l < NC; l++) { tot += fix_label[l]; aver_res_label[l] = (float)fixval(fix_label[l], ls); }
            aver_res = (float)fixval(tot, ls) / float(2 * N);
            res_sq = quant ? std::ldexp((double)rs_q, -2 * rexp - ORC_QFRAC_BITS) : rs_d;
        } else {
            float tot = 0.f;
            for (int l = 0; l < NC; l++) tot += aver_res_label[l];
            aver_res = tot / float(2 * N);
            res_sq = rs_f;
        }
        for (int l = 0; l < NC; l++) aver_res_label[l] /= float(2 * num_pix_label[l]);

        if (p.enable_segmentation) { StageTimer tm(&stage_s[5], &stage_s[4]); solveSegmIteration(aver_res_label, aver_res_old, lap); } /* :672 */

        float delta_sol_max = 0.f; /* :676 */
        for (int i = 0; i < 6; i++) delta_sol_max = std::max(delta_sol_max, std::fabs(prev_sol[i] - Var[i]));
        for (int i = 0; i < 6; i++) prev_sol[i] = Var[i];
        it_done = k;
        total_irls++;
        if (k <= ORC_TRACE_MAX_IRLS) {
            float* ti = tr + ORC_TRACE_HDR + (k - 1) * ORC_TRACE_IRLS;
            for (int i = 0; i < 6; i++) ti[i] = Var[i];
            for (int l = 0; l < NC; l++) ti[6 + l] = b_segm[l];
            ti[30] = aver_res; ti[31] = delta_sol_max; ti[32] = (float)res_sq;
        }
        if ((delta_sol_max < p.irls_delta_threshold) || (k == p.max_iter_irls) || !(aver_res > 0.f)) break;
    }
    tr[4] = (float)it_done;

    /* :689 est_cov = AtA.inverse() * res.squaredNorm() */
    if (exact) {
        double F[36]; unsigned char zero[6];
        for (int i = 0; i < 36; i++) F[i] = AtA[i];
        ldlt_factor<double>(6, F, zero);
        for (int c = 0; c < 6; c++) {
            double e[6] = {0, 0, 0, 0, 0, 0}, x[6];
            e[c] = 1;
            ldlt_solve_factored<double>(6, F, zero, e, x);
            for (int r = 0; r < 6; r++) est_cov[r * 6 + c] = x[r] * res_sq;
        }
    } else {
        float F[36], inv[36];
        for (int i = 0; i < 36; i++) F[i] = (float)AtA[i];
        general_inverse<float>(6, F, inv);
        for (int i = 0; i < 36; i++) est_cov[i] = (double)(inv[i] * (float)res_sq);
    }
    { StageTimer tm(&stage_s[6], &stage_s[4]); filterEstimateAndComputeT(Var); }
}

/* FrontEnd.cpp:713-772 */
template <class T>
static void filterAndCompose(orc_ctx& c, float* twist) {
    T tw[6];
    for (int i = 0; i < 6; i++) tw[i] = (T)twist[i];
    T Tod[16];
    for (int i = 0; i < 16; i++) Tod[i] = (T)c.T_odometry[i];
    if (c.p.use_motion_filter) {
        T cov[36], ev[6], V[36];
        for (int i = 0; i < 36; i++) cov[i] = (T)c.est_cov[i];
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < i; j++) { const T m = (T)0.5 * (cov[i * 6 + j] + cov[j * 6 + i]); cov[i * 6 + j] = m; cov[j * 6 + i] = m; }
        jacobi_eig6<T>(cov, ev, V);
        T kai_b[6], kai_b_old[6], kai_loc_sub[6], lg[6];
        se3_log<T>(Tod, lg); /* :736-738: old twist minus the motion already applied at coarser levels */
        for (int i = 0; i < 6; i++) kai_loc_sub[i] = (T)c.twist_odometry_old[i] - lg[i];
        for (int i = 0; i < 6; i++) { /* Bii orthogonal: solve(Bii, x) = Bii^T x (:729,741) */
            T a = 0, b = 0;
            for (int k = 0; k < 6; k++) { a += V[k * 6 + i] * tw[k]; b += V[k * 6 + i] * kai_loc_sub[k]; }
            kai_b[i] = a; kai_b_old[i] = b;
        }
        /* expf(-level) evaluated once on the host in float (FrontEnd.cpp:745) */
        const float e = expf(-(float)c.level);
        const T cf = (T)(c.p.previous_speed_eig_weight * e), df = (T)(c.p.previous_speed_const_weight * e);
        T fil[6];
        for (int i = 0; i < 6; i++) fil[i] = (kai_b[i] + (cf * ev[i] + df) * kai_b_old[i]) / ((T)1 + cf * ev[i] + df);
        for (int i = 0; i < 6; i++) { /* :755 twist = Bii * kai_b_fil */
            T a = 0;
            for (int k = 0; k < 6; k++) a += V[i * 6 + k] * fil[k];
            tw[i] = a;
        }
    }
    for (int i = 0; i < 6; i++) { twist[i] = (float)tw[i]; c.twist_level_odometry[i] = twist[i]; }
    for (int i = 0; i < 6; i++) tw[i] = (T)twist[i];
    T E[16], Tn[16];
    se3_exp<T>(tw, E);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            T s = 0;
            for (int k = 0; k < 4; k++) s += E[i * 4 + k] * Tod[k * 4 + j];
            Tn[i * 4 + j] = s;
        }
    for (int i = 0; i < 16; i++) c.T_odometry[i] = (float)Tn[i];
    for (int i = 0; i < 16; i++) Tn[i] = (T)c.T_odometry[i];
    T lg[6];
    se3_log<T>(Tn, lg);
    for (int i = 0; i < 6; i++) c.twist_odometry[i] = (float)lg[i];
}
/* reference-literal policy: float throughout, the general solves of the source (colPivHouseholderQr / inverse,
 * FrontEnd.cpp:729,741,755) as Gaussian elimination with partial pivoting */
static void filterAndComposeLiteral(orc_ctx& c, float* twist) {
    float Tod[16];
    for (int i = 0; i < 16; i++) Tod[i] = c.T_odometry[i];
    if (c.p.use_motion_filter) {
        float cov[36], ev[6], V[36];
        for (int i = 0; i < 36; i++) cov[i] = (float)c.est_cov[i];
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < i; j++) { const float m = 0.5f * (cov[i * 6 + j] + cov[j * 6 + i]); cov[i * 6 + j] = m; cov[j * 6 + i] = m; }
        jacobi_eig6<float>(cov, ev, V);  /* Bii = V */
        float M[36], kai_b[6], kai_b_old[6], lg[6];
        for (int i = 0; i < 36; i++) M[i] = V[i];
        for (int i = 0; i < 6; i++) kai_b[i] = twist[i];
        gauss_solve<float>(6, 1, M, kai_b);                         /* :729 */
        se3_log<float>(Tod, lg);                                    /* :736-738 */
        for (int i = 0; i < 6; i++) kai_b_old[i] = c.twist_odometry_old[i] - lg[i];
        for (int i = 0; i < 36; i++) M[i] = V[i];
        gauss_solve<float>(6, 1, M, kai_b_old);                     /* :741 */
        const float e = expf(-(float)c.level);
        const float cf = c.p.previous_speed_eig_weight * e, df = c.p.previous_speed_const_weight * e;  /* :745 */
        float fil[6];
        for (int i = 0; i < 6; i++) fil[i] = (kai_b[i] + (cf * ev[i] + df) * kai_b_old[i]) / (1.f + cf * ev[i] + df);  /* :750 */
        float Vinv[36];
        general_inverse<float>(6, V, Vinv);
        gauss_solve<float>(6, 1, Vinv, fil);                        /* :755 twist = Bii.inverse().colPivHouseholderQr().solve(kai_b_fil) */
        for (int i = 0; i < 6; i++) twist[i] = fil[i];
    }
    for (int i = 0; i < 6; i++) c.twist_level_odometry[i] = twist[i];
    float E[16], Tn[16], lg[6];
    se3_exp<float>(twist, E);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float s = 0.f;
            for (int k = 0; k < 4; k++) s += E[i * 4 + k] * Tod[k * 4 + j];
            Tn[i * 4 + j] = s;
        }
    for (int i = 0; i < 16; i++) c.T_odometry[i] = Tn[i];
    se3_log<float>(Tn, lg);
    for (int i = 0; i < 6; i++) c.twist_odometry[i] = lg[i];
}
void orc_ctx::filterEstimateAndComputeT(float* twist) {
    if (accum != ORC_ACCUM_F32) filterAndCompose<double>(*this, twist);
    else filterAndComposeLiteral(*this, twist);
}

/* FrontEnd.cpp:1071-1146 */
void orc_ctx::runSolver(bool create_image_pyr, int stop_step) {
    status = 0; total_irls = 0;
    std::fill(trace.begin(), trace.end(), 0.f);
    if (create_image_pyr) createImagePyramid(false);
    if (p.enable_segmentation) {
        StageTimer tm(&stage_s[1]);
        kMeans3DCoord();
        createClustersPyramidUsingKMeans();
    } else {
        for (auto& L : clusterAllocation) L.fill(0);
    }
    for (int i = 0; i < 16; i++) T_odometry[i] = (i % 5 == 0) ? 1.f : 0.f;
    for (int i = 0; i < 6; i++) twist_odometry[i] = 0.f;

    for (int i = 0; i < p.ctf_levels; i++)
        for (int k = 0; k < p.max_iter_per_level; k++) {
            level = i;
            const int s = 1 << (p.ctf_levels - (i + 1));
            cols_i = cols / s; rows_i = rows / s;
            image_level = p.ctf_levels - i - 1;
            const int step = i * p.max_iter_per_level + k;
            cur_trace = &trace[(size_t)step * ORC_TRACE_STEP];
            cur_trace[0] = 1.f; cur_trace[1] = (float)i; cur_trace[2] = (float)k;
            if ((i == 0) && (k == 0)) { /* :1103-1110 */
                depthWarpedPyr[image_level] = depthPredPyr[image_level];
                intensityWarpedPyr[image_level] = intensityPredPyr[image_level];
                xxWarpedPyr[image_level] = xxPredPyr[image_level];
                yyWarpedPyr[image_level] = yyPredPyr[image_level];
            } else {
                StageTimer tm(&stage_s[2]);
                warpImagesAccurateInverse();
            }
            {
                StageTimer tm(&stage_s[3]);
                calculateCoord();
                calculateDerivatives();
                computeWeights();
                computeSegPrior();
            }
            for (int l = 0; l < NC; l++) { cur_trace[8 + l] = b_prior[l]; cur_trace[32 + l] = lambda_t_w[l]; }
            if (step == stop_step) return;
            {
                StageTimer tm(&stage_s[4]);
                solveOdometryAndSegmJoint();
            }
            for (int q = 0; q < 6; q++) { cur_trace[56 + q] = twist_level_odometry[q]; cur_trace[79 + q] = twist_odometry[q]; }
            for (int q = 0; q < 16; q++) cur_trace[62 + q] = T_odometry[q];
            cur_trace[78] = (float)status;
            /* :1130 */
            if (accum == ORC_ACCUM_F32) {
                float nf = 0.f;
                for (int q = 0; q < 6; q++) nf += twist_level_odometry[q] * twist_level_odometry[q];
                if (std::sqrt(nf) < p.outer_exit_threshold) break;
            } else {
                double nrm = 0;
                for (int q = 0; q < 6; q++) nrm += (double)twist_level_odometry[q] * (double)twist_level_odometry[q];
                if (std::sqrt(nrm) < (double)p.outer_exit_threshold) break;
            }
        }
    /* :1139-1144: twist_odometry_old = R^-1 * twist_odometry on each 3-half, R^-1 formed in double (MRPT) then cast */
    double R[9], Ri[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[i * 3 + j] = (double)T_odometry[i * 4 + j];
    if (accum == ORC_ACCUM_F32) general_inverse<double>(3, R, Ri);
    const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
    const double id = 1.0 / det;
    if (accum != ORC_ACCUM_F32) {
    Ri[0] = (R[4] * R[8] - R[5] * R[7]) * id; Ri[1] = (R[2] * R[7] - R[1] * R[8]) * id; Ri[2] = (R[1] * R[5] - R[2] * R[4]) * id;
    Ri[3] = (R[5] * R[6] - R[3] * R[8]) * id; Ri[4] = (R[0] * R[8] - R[2] * R[6]) * id; Ri[5] = (R[2] * R[3] - R[0] * R[5]) * id;
    Ri[6] = (R[3] * R[7] - R[4] * R[6]) * id; Ri[7] = (R[1] * R[6] - R[0] * R[7]) * id; Ri[8] = (R[0] * R[4] - R[1] * R[3]) * id;
    }
    for (int h = 0; h < 2; h++)
        for (int i = 0; i < 3; i++) {
            float s = 0.f;
            for (int j = 0; j < 3; j++) s += (float)Ri[i * 3 + j] * twist_odometry[3 * h + j];
            twist_odometry_old[3 * h + i] = s;
        }
}

/* SegmentationBackground.cpp:176-197 */
void orc_ctx::buildSegmImage() {
    StageTimer tm(&stage_s[7]);
    const ImgI& labels_maxres = clusterAllocation[0];
    for (int u = 0; u < cols; u++)
        for (int v = 0; v < rows; v++) {
            if (labels_maxres(v, u) == NC) { b_segm_perpixel(v, u) = 1.f; continue; }
            b_segm_perpixel(v, u) = std::max(0.f, std::min(1.f, b_segm[labels_maxres(v, u)]));
            if (perClusterAverageResidual[labels_maxres(v, u)] < 0.017) /* NaN unless the 5-frame history ran */
                b_segm_perpixel(v, u) = std::max(b_segm_perpixel(v, u), 1.0f - b_segm_perpixel(v, u));
        }
}

/* FrontEnd.cpp:896-1069: warp the frame of five frames ago into the current view with the composed increments and
 * average |dz| + k|dI| per cluster of the current frame.  `index` is the driver's im_count. */
void orc_ctx::computeResidualsAgainstPreviousImage(int index) {
    const int idx_to_warp = (index - bufferLength) % bufferLength;
    const int trans_start = index - bufferLength + 1;
    const bool exact = (accum != ORC_ACCUM_F32);
    /* :901-909  T = (prod odomBuffer * T_odometry)^-1 */
    float T[16];
    if (!exact) {
        float M[16], tmp[16];
        for (int i = 0; i < 16; i++) M[i] = (i % 5 == 0) ? 1.f : 0.f;
        auto mul = [&](const float* B) {
            for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
                float acc = 0.f;
                for (int k = 0; k < 4; k++) acc += M[i * 4 + k] * B[k * 4 + j];
                tmp[i * 4 + j] = acc;
            }
            for (int i = 0; i < 16; i++) M[i] = tmp[i];
        };
        for (int i = trans_start; i < index; i++) mul(odomBuffer[i % bufferLength]);
        mul(T_odometry);
        general_inverse<float>(4, M, T);
    } else { /* products in double of the float increments, one rounding, rigid inverse */
        double M[16], tmp[16];
        for (int i = 0; i < 16; i++) M[i] = (i % 5 == 0) ? 1.0 : 0.0;
        auto mul = [&](const float* B) {
            for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
                double acc = 0.0;
                for (int k = 0; k < 4; k++) acc += M[i * 4 + k] * (double)B[k * 4 + j];
                tmp[i * 4 + j] = acc;
            }
            for (int i = 0; i < 16; i++) M[i] = tmp[i];
        };
        for (int i = trans_start; i < index; i++) mul(odomBuffer[i % bufferLength]);
        mul(T_odometry);
        float Mf[16];
        for (int i = 0; i < 16; i++) Mf[i] = (float)M[i];
        rigid_inverse<double>(Mf, T);
    }

    const float inv_f_i = 2.f * std::tan(0.5f * p.fovh) / float(cols);
    const float disp_u_i = 0.5f * float(cols - 1);
    const float disp_v_i = 0.5f * float(rows - 1);
    const float f = float(cols) / (2.f * std::tan(0.5f * p.fovh));
    const Img& dref = depthBuffer[idx_to_warp];
    const Img& iref = intensityBuffer[idx_to_warp];
    depthWarpedRefference.fill(0.f);
    intensityWarpedRefference.fill(0.f);
    Img intensity_diff = intensityCurrent;
    const ImgI& labels_ref = clusterAllocation[0];
    const size_t np = (size_t)rows * cols;
    std::vector<float> wacu(np, 0.f);
    std::vector<int64_t> dfix, ifix;
    std::vector<int32_t> wfix;
    if (exact) { dfix.assign(np, 0); ifix.assign(np, 0); wfix.assign(np, 0); }
    const int cols_lim = 100 * (cols - 1);
    const int rows_lim = 100 * (rows - 1);

    auto splat = [&](int v, int u, int w, float depth_w, float intensity_w) {
        const size_t k = (size_t)v * cols + u;
        if (exact) {
            dfix[k] += (int64_t)w * fixq(depth_w, 32);
            ifix[k] += (int64_t)w * fixq(intensity_w, 22);
            wfix[k] += w;
        } else {
            depthWarpedRefference.a[k] += float(w) * depth_w;
            intensityWarpedRefference.a[k] += float(w) * intensity_w;
            wacu[k] += float(w);
        }
    };

    for (int j = 0; j < cols; j++)
        for (int i = 0; i < rows; i++) {
            const float z = dref(i, j);
            if (z != 0.f && depthCurrent(i, j) != 0.f) { /* :951 tests the CURRENT depth at the source pixel */
                const float intensity_w = iref(i, j);
                const float xr = (inv_f_i * (float(j) - disp_u_i)) * z; /* :925-929 */
                const float yr = (inv_f_i * (float(i) - disp_v_i)) * z;
                const float x_w = T[0] * xr + T[1] * yr + T[2] * z + T[3];
                const float y_w = T[4] * xr + T[5] * yr + T[6] * z + T[7];
                const float depth_w = T[8] * xr + T[9] * yr + T[10] * z + T[11];
                const float fu = 100.f * (f * x_w / depth_w + disp_u_i);
                const float fv = 100.f * (f * y_w / depth_w + disp_v_i);
                if (!(std::fabs(fu) < 1.0e9f) || !(std::fabs(fv) < 1.0e9f)) continue; /* see warpImagesAccurateInverse */
                const int uwarp = int(fu);
                const int vwarp = int(fv);
                if ((uwarp >= 0) && (uwarp < cols_lim) && (vwarp >= 0) && (vwarp < rows_lim)) {
                    const int uwarp_l = uwarp - uwarp % 100;
                    const int uwarp_r = uwarp_l + 100;
                    const int vwarp_d = vwarp - vwarp % 100;
                    const int vwarp_u = vwarp_d + 100;
                    const int delta_r = uwarp_r - uwarp;
                    const int delta_l = 100 - delta_r;
                    const int delta_u = vwarp_u - vwarp;
                    const int delta_d = 100 - delta_u;
                    if (std::min(delta_r, delta_l) + std::min(delta_u, delta_d) < 5) {
                        const int ind_u = delta_r > delta_l ? uwarp_l / 100 : uwarp_r / 100;
                        const int ind_v = delta_u > delta_d ? vwarp_d / 100 : vwarp_u / 100;
                        splat(ind_v, ind_u, 200, depth_w, intensity_w);
                    } else {
                        const int v_d = vwarp_d / 100, u_l = uwarp_l / 100;
                        const int v_u = v_d + 1, u_r = u_l + 1;
                        splat(v_u, u_r, delta_l + delta_d, depth_w, intensity_w);
                        splat(v_u, u_l, delta_r + delta_d, depth_w, intensity_w);
                        splat(v_d, u_r, delta_l + delta_u, depth_w, intensity_w);
                        splat(v_d, u_l, delta_r + delta_u, depth_w, intensity_w);
                    }
                }
            } else {
                intensity_diff(i, j) = 0.f;
            }
        }

    /* :1025-1035 */
    for (size_t k = 0; k < np; k++) {
        if (exact) {
            if (wfix[k] != 0) {
                intensityWarpedRefference.a[k] = (float)((double)ifix[k] / ((double)wfix[k] * 4194304.0));
                depthWarpedRefference.a[k] = (float)((double)dfix[k] / ((double)wfix[k] * 4294967296.0));
            }
        } else if (wacu[k] != 0.f) {
            intensityWarpedRefference.a[k] /= wacu[k];
            depthWarpedRefference.a[k] /= wacu[k];
        }
    }

    /* :1039-1068 */
    float sum_f[NC];
    int64_t sum_q[NC];
    int num_pix_label[NC];
    for (int l = 0; l < NC; l++) { sum_f[l] = std::numeric_limits<float>::quiet_NaN(); sum_q[l] = 0; num_pix_label[l] = 1; }
    for (int j = 0; j < cols; j++)
        for (int i = 0; i < rows; i++) {
            const float dr = depthCurrent(i, j) - depthWarpedRefference(i, j);
            const float ir = intensity_diff(i, j) - intensityWarpedRefference(i, j);
            const float cr = std::fabs(dr) + p.k_photometric_res * std::fabs(ir);
            cumulativeResiduals(i, j) = cr;
            if (depthWarpedRefference(i, j) != 0.f && depthCurrent(i, j) != 0.f) {
                const int l = labels_ref(i, j);
                if (exact) sum_q[l] += fixq(cr, 32);
                else if (std::isnan(sum_f[l])) sum_f[l] = cr;
                else sum_f[l] += cr;
                num_pix_label[l]++;
            }
        }
    for (int l = 0; l < NC; l++) {
        if (exact) sum_f[l] = (num_pix_label[l] > 1) ? (float)fixval(sum_q[l], 32) : std::numeric_limits<float>::quiet_NaN();
        perClusterAverageResidual[l] = sum_f[l] / float(2 * num_pix_label[l]);
    }
}

/* =====================================================================================
 * Depth pre-filter: the step right before the path
 * ===================================================================================== */
/* exp restated with IEEE float operations only (Cody-Waite reduction by ln2, degree-7 Horner, exact scaling), so that
 * the CPU and CUDA evaluations agree bit for bit; < 2 ulp from expf over the range the filter uses. */
static inline float det_expf(float a) {
    if (!(a > -87.f)) return 0.f;
    if (a > 0.f) a = 0.f; /* the filter's argument is -(non-negative) */
    const float k = rintf(a * 1.44269504088896341f);
    float r = a - k * 0.693359375f;      /* ln2 high part: 10 significant bits, k*hi is exact */
    r = r - k * -2.12194440e-4f;         /* ln2 low part */
    float p = 1.f / 5040.f;
    p = p * r + 1.f / 720.f;
    p = p * r + 1.f / 120.f;
    p = p * r + 1.f / 24.f;
    p = p * r + 1.f / 6.f;
    p = p * r + 0.5f;
    p = p * r + 1.f;
    p = p * r + 1.f;
    return ldexpf(p, (int)k);
}

/* Shaders/depth_bilateral.frag:30-76 followed by Shaders/depth_metric.frag:28-40, as Reconstruction::getFilteredDepth
 * chains them (Reconstruction.cpp:722-732): 13x13 bilateral on raw u16 millimetres, invalid outside [300, maxD*1000],
 * rounded back to u16, then metres.  Texels are addressed ideally (texture(cx/cols, cy/rows) -> texel (cx, cy)).
 * exact == 0: libm expf (closest to what a GLSL implementation does); exact != 0: det_expf, the CUDA contract.
 * PARITY UNPINNED for this function: the reference runs it as a GLSL fragment shader whose exp/round/texel addressing
 * are implementation-defined and there is no GL context here; both policies are restatements of the shader text. */
static void filter_depth(const uint16_t* in, int rows, int cols, float maxD, int exact, float* out) {
    const float sigma_space2_inv_half = 0.024691358f;
    const float sigma_color2_inv_half = 0.000555556f;
    const int R = 6, D = R * 2 + 1;
    const unsigned lim = (unsigned)(maxD * 1000.0f);
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < cols; x++) {
            const unsigned value = in[(size_t)y * cols + x];
            unsigned filtered = 0;
            if (!(value > lim || value < 300u)) {
                const int tx = std::min(x - D / 2 + D, cols);
                const int ty = std::min(y - D / 2 + D, rows);
                float sum1 = 0.f, sum2 = 0.f;
                for (int cy = std::max(y - D / 2, 0); cy < ty; ++cy)
                    for (int cx = std::max(x - D / 2, 0); cx < tx; ++cx) {
                        const unsigned tmp = in[(size_t)cy * cols + cx];
                        const float space2 = (float(x) - float(cx)) * (float(x) - float(cx)) + (float(y) - float(cy)) * (float(y) - float(cy));
                        const float color2 = (float(value) - float(tmp)) * (float(value) - float(tmp));
                        const float arg = -(space2 * sigma_space2_inv_half + color2 * sigma_color2_inv_half);
                        const float weight = exact ? det_expf(arg) : expf(arg);
                        sum1 += float(tmp) * weight;
                        sum2 += weight;
                    }
                filtered = (unsigned)roundf(sum1 / sum2);
            }
            out[(size_t)y * cols + x] = (filtered > lim || filtered < 300u) ? 0.f : float(filtered) / 1000.0f;
        }
}

/* =====================================================================================
 * C ABI
 * ===================================================================================== */
extern "C" {

orc_ctx* orc_create(const orc_params* p, int accum_mode) {
    if (!p || p->rows <= 0 || p->cols <= 0 || p->ctf_levels < 1) return nullptr;
    if (p->enable_segmentation && p->ctf_levels < 2) return nullptr;
    orc_ctx* c = new orc_ctx();
    c->init(*p, accum_mode);
    return c;
}
void orc_destroy(orc_ctx* c) { delete c; }
void orc_set_params(orc_ctx* c, const orc_params* p) {
    /* only scalar tuning may change (kb is rewritten per frame, StaticFusion-datasets.cpp:156-165) */
    const int r = c->p.rows, co = c->p.cols, L = c->p.ctf_levels, m = c->p.max_iter_per_level;
    c->p = *p;
    c->p.rows = r; c->p.cols = co; c->p.ctf_levels = L; c->p.max_iter_per_level = m;
}
void orc_set_current(orc_ctx* c, const float* depth, const float* intensity) {
    std::memcpy(c->depthCurrent.a.data(), depth, sizeof(float) * c->depthCurrent.a.size());
    std::memcpy(c->intensityCurrent.a.data(), intensity, sizeof(float) * c->intensityCurrent.a.size());
}
void orc_set_prediction(orc_ctx* c, const float* depth, const float* intensity) {
    std::memcpy(c->depthPrediction.a.data(), depth, sizeof(float) * c->depthPrediction.a.size());
    std::memcpy(c->intensityPrediction.a.data(), intensity, sizeof(float) * c->intensityPrediction.a.size());
}
void orc_set_twist_old(orc_ctx* c, const float t[6]) { for (int i = 0; i < 6; i++) c->twist_odometry_old[i] = t[i]; }
void orc_create_image_pyramid(orc_ctx* c, int old_im) { c->createImagePyramid(old_im != 0); }
void orc_run_solver(orc_ctx* c, int create_image_pyr, int stop_step) { c->runSolver(create_image_pyr != 0, stop_step); }
void orc_build_segm_image(orc_ctx* c) { c->buildSegmImage(); }
void orc_kmeans(orc_ctx* c) { c->kMeans3DCoord(); c->createClustersPyramidUsingKMeans(); }
void orc_warp_level(orc_ctx* c, int image_level, const float T[16]) {
    for (int i = 0; i < 16; i++) c->T_odometry[i] = T[i];
    c->image_level = image_level;
    c->rows_i = c->rows >> image_level; c->cols_i = c->cols >> image_level;
    c->warpImagesAccurateInverse();
}
void orc_get_T(const orc_ctx* c, float out[16]) { for (int i = 0; i < 16; i++) out[i] = c->T_odometry[i]; }
void orc_get_twists(const orc_ctx* c, float a[6], float b[6], float l[6]) {
    for (int i = 0; i < 6; i++) { if (a) a[i] = c->twist_odometry[i]; if (b) b[i] = c->twist_odometry_old[i]; if (l) l[i] = c->twist_level_odometry[i]; }
}
void orc_get_b_segm(const orc_ctx* c, float out[NC]) { for (int l = 0; l < NC; l++) out[l] = c->b_segm[l]; }
void orc_get_b_perpixel(const orc_ctx* c, float* out) { std::memcpy(out, c->b_segm_perpixel.a.data(), sizeof(float) * c->b_segm_perpixel.a.size()); }
int orc_get_labels(const orc_ctx* c, int L, int32_t* out) {
    if (L < 0 || L >= c->pyr_levels) return -1;
    std::memcpy(out, c->clusterAllocation[L].a.data(), sizeof(int32_t) * c->clusterAllocation[L].a.size());
    return 0;
}
void orc_get_kmeans(const orc_ctx* c, float out[3 * NC]) {
    for (int r = 0; r < 3; r++) for (int l = 0; l < NC; l++) out[r * NC + l] = c->kmeans[r][l];
}
void orc_get_connectivity(const orc_ctx* c, uint8_t out[NC * NC]) {
    for (int i = 0; i < NC; i++) for (int j = 0; j < NC; j++) out[i * NC + j] = c->connectivity[i][j] ? 1 : 0;
}
/* ring buffers of the drivers (StaticFusion-datasets.cpp:114-116, 182-184) */
void orc_buffer_set(orc_ctx* c, int slot, const float* depth, const float* intensity, const float T[16]) {
    const int b = ((slot % orc_ctx::bufferLength) + orc_ctx::bufferLength) % orc_ctx::bufferLength;
    std::memcpy(c->depthBuffer[b].a.data(), depth, sizeof(float) * c->depthBuffer[b].a.size());
    std::memcpy(c->intensityBuffer[b].a.data(), intensity, sizeof(float) * c->intensityBuffer[b].a.size());
    for (int i = 0; i < 16; i++) c->odomBuffer[b][i] = T[i];
}
void orc_buffer_push(orc_ctx* c, int index) { /* depthCurrent, intensityCurrent, T_odometry -> slot index % 5 */
    orc_buffer_set(c, index, c->depthCurrent.a.data(), c->intensityCurrent.a.data(), c->T_odometry);
}
void orc_compute_residuals_against_previous_image(orc_ctx* c, int index) { c->computeResidualsAgainstPreviousImage(index); }
void orc_get_per_cluster_average_residual(const orc_ctx* c, float out[NC]) { for (int l = 0; l < NC; l++) out[l] = c->perClusterAverageResidual[l]; }
void orc_set_per_cluster_average_residual(orc_ctx* c, const float in[NC]) { for (int l = 0; l < NC; l++) c->perClusterAverageResidual[l] = in[l]; }
void orc_set_T(orc_ctx* c, const float T[16]) { for (int i = 0; i < 16; i++) c->T_odometry[i] = T[i]; }
int orc_get_residual_image(const orc_ctx* c, const char* name, float* out) {
    const std::string n(name);
    const Img* src = nullptr;
    if (n == "depth_warped_ref") src = &c->depthWarpedRefference;
    else if (n == "intensity_warped_ref") src = &c->intensityWarpedRefference;
    else if (n == "cumulative") src = &c->cumulativeResiduals;
    if (!src) return -2;
    std::memcpy(out, src->a.data(), sizeof(float) * src->a.size());
    return 0;
}
void orc_filter_depth(const uint16_t* depth_mm, int rows, int cols, float max_depth_m, int exact, float* out) {
    filter_depth(depth_mm, rows, cols, max_depth_m, exact, out);
}
float orc_det_expf(float a) { return det_expf(a); }

/* StaticFusion::loadImageFromSequenceAssoc, FrontEnd.cpp:216-254 (conversion half) */
void orc_convert_frame(const uint8_t* bgr, const uint16_t* depth_raw, int full_rows, int full_cols, int res_factor,
                       float* intensity, float* depth, uint16_t* depth_mm, uint8_t* color) {
    const int height = full_rows / res_factor, width = full_cols / res_factor;
    const float norm_factor = 1.f / 255.f;                                               /* :218 */
    const float mm_to_m = (float)(1.0 / 1000.0);                                         /* :243 convertTo(CV_32FC1, 1.0/1000.0): float(src) * float(alpha) */
    for (int v = 0; v < height; v++)
        for (int u = 0; u < width; u++) {
            const size_t src = (size_t)(height * res_factor - res_factor * v - 1) * full_cols + (size_t)res_factor * u;  /* :231 vertical flip + decimation */
            const size_t dst = (size_t)v * width + u;
            const float r = norm_factor * (float)bgr[3 * src + 0];                       /* :232-234: channel 0 of a BGR image is named r */
            const float g = norm_factor * (float)bgr[3 * src + 1];
            const float b = norm_factor * (float)bgr[3 * src + 2];
            if (intensity) intensity[dst] = 0.299f * r + 0.587f * g + 0.114f * b;        /* :236, left to right */
            if (color) {                                                                 /* :237 Vec3b(r*255, g*255, b*255): float -> uchar truncation */
                color[3 * dst + 0] = (uint8_t)(r * 255); color[3 * dst + 1] = (uint8_t)(g * 255); color[3 * dst + 2] = (uint8_t)(b * 255);
            }
            if (depth) depth[dst] = (float)depth_raw[src] * mm_to_m;                     /* :249 */
            if (depth_mm) depth_mm[dst] = depth_raw[src];                                /* :244,250 convertTo(CV_16U, 1.0) is a copy */
        }
}

/* Reconstruction.cpp:256,265: currPose = currPose * (*inPose), Eigen Matrix4f product (sequential k = 0..3 float sums) */
void orc_pose_compose(const float A[16], const float B[16], float out[16]) {
    float r[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float s = A[i * 4 + 0] * B[0 * 4 + j];
            for (int k = 1; k < 4; k++) s += A[i * 4 + k] * B[k * 4 + j];
            r[i * 4 + j] = s;
        }
    std::memcpy(out, r, sizeof(r));
}

/* Eigen::Quaternionf(const Matrix3f&) as used by Datasets.cpp:259 and Reconstruction.cpp:480 (Eigen 3 Quaternion.h,
 * quaternionbase_assign_impl<Other,3,3>): not in the reference tree, published algorithm restated */
void orc_quat_from_rotation(const float T[16], float q[4]) {
    auto m = [&](int r, int c) { return T[r * 4 + c]; };
    float t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0.f) {
        t = std::sqrt(t + 1.f);
        q[3] = 0.5f * t;
        t = 0.5f / t;
        q[0] = (m(2, 1) - m(1, 2)) * t; q[1] = (m(0, 2) - m(2, 0)) * t; q[2] = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.f);
        q[i] = 0.5f * t;
        t = 0.5f / t;
        q[3] = (m(k, j) - m(j, k)) * t; q[j] = (m(j, i) + m(i, j)) * t; q[k] = (m(k, i) + m(i, k)) * t;
    }
}
int orc_get_status(const orc_ctx* c) { return c->status; }
void orc_get_stage_seconds(orc_ctx* c, double out[ORC_NUM_STAGES], int reset) {
    for (int i = 0; i < ORC_NUM_STAGES; i++) { out[i] = c->stage_s[i]; if (reset) c->stage_s[i] = 0; }
}
int orc_get_total_irls(const orc_ctx* c) { return c->total_irls; }

int orc_get_image(const orc_ctx* c, const char* name, int L, float* out) {
    if (L < 0 || L >= c->pyr_levels) return -1;
    const std::string n(name);
    const int r = c->rows >> L, co = c->cols >> L;
    const Img* src = nullptr;
    if (n == "depth") src = &c->depthPyr[L];
    else if (n == "intensity") src = &c->intensityPyr[L];
    else if (n == "xx") src = &c->xxPyr[L];
    else if (n == "yy") src = &c->yyPyr[L];
    else if (n == "depth_pred") src = &c->depthPredPyr[L];
    else if (n == "intensity_pred") src = &c->intensityPredPyr[L];
    else if (n == "depth_warped") src = &c->depthWarpedPyr[L];
    else if (n == "intensity_warped") src = &c->intensityWarpedPyr[L];
    else if (n == "depth_inter") src = &c->depthInterPyr[L];
    else if (n == "intensity_inter") src = &c->intensityInterPyr[L];
    else if (n == "xx_inter") src = &c->xxInterPyr[L];
    else if (n == "yy_inter") src = &c->yyInterPyr[L];
    if (src) { std::memcpy(out, src->a.data(), sizeof(float) * (size_t)r * co); return 0; }
    /* full-res holders with the active level in the top-left block */
    const Img* big = nullptr;
    if (n == "dcu") big = &c->dcu; else if (n == "dcv") big = &c->dcv; else if (n == "dct") big = &c->dct;
    else if (n == "ddu") big = &c->ddu; else if (n == "ddv") big = &c->ddv; else if (n == "ddt") big = &c->ddt;
    else if (n == "weights_c") big = &c->weights_c; else if (n == "weights_d") big = &c->weights_d;
    if (big) {
        for (int v = 0; v < r; v++) for (int u = 0; u < co; u++) out[(size_t)v * co + u] = (*big)(v, u);
        return 0;
    }
    if (n == "null") {
        for (size_t i = 0; i < (size_t)r * co; i++) out[i] = c->Null[i] ? 1.f : 0.f;
        return 0;
    }
    return -2;
}
int orc_trace_size(const orc_ctx* c) { return (int)c->trace.size(); }
void orc_get_trace(const orc_ctx* c, float* out) { std::memcpy(out, c->trace.data(), sizeof(float) * c->trace.size()); }

void orc_se3_exp(const double xi[6], double T[16]) { se3_exp<double>(xi, T); }
void orc_se3_log(const double T[16], double xi[6]) { se3_log<double>(T, xi); }
int orc_ldlt_solve(int n, const double* A, const double* b, double* x) {
    std::vector<double> F(A, A + (size_t)n * n);
    std::vector<unsigned char> zero(n);
    const int nz = ldlt_factor<double>(n, F.data(), zero.data());
    ldlt_solve_factored<double>(n, F.data(), zero.data(), b, x);
    return nz;
}
void orc_jacobi_eig6(const double A[36], double ev[6], double V[36]) { jacobi_eig6<double>(A, ev, V); }

}  // extern "C"
