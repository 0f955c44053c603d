// sf_device.cuh — device-side state and small dense algebra of the B200 StaticFusion solver.
//
// Numerics contract: every per-pixel expression is evaluated in float32
// with the operation order of the reference source and no FMA contraction (the library is
// compiled with -fmad=false; fused adds are written explicitly where the contract allows).
// Every cross-pixel sum is an order-independent fixed-point (integer) sum - the normal equations
// through double-precision FMAs onto accumulators whose bit patterns are those integers; the 6x6 / 24x24 solves, the 6x6 eigen-decomposition and SE(3) exp/log run in
// double on one thread / one warp per pair.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/staticfusion_b200.h"

namespace sf {

constexpr int NC = SF_NUM_CLUSTERS;
constexpr int MAX_LEVELS = 8;
constexpr int NPLANES = 11;  // debug planes: d, x, y, dcu, dcv, dct, ddu, ddv, ddt, wc, wd (only kept when tracing)
enum Plane { PL_D = 0, PL_X, PL_Y, PL_DCU, PL_DCV, PL_DCT, PL_DDU, PL_DDV, PL_DDT, PL_WC, PL_WD };
// the two Jacobian rows of a pixel built with the RAW pre-weights (FrontEnd.cpp:552-585): what the IRLS passes stream
constexpr int NROWPL = 14;   // colour row a0..a5, colour rhs b, depth row a0..a5, depth rhs b
constexpr int RW_AC = 0, RW_BC = 6, RW_AD = 7, RW_BD = 13;
constexpr uint8_t LABEL_NONE = NC;      // no depth (clusterAllocation value 24, KMeans.cpp:70)
constexpr uint8_t VLABEL_INVALID = 255; // pixel not in validPixels (FrontEnd.cpp:417)

// Rows are stored as TILES: one 64-pixel tile = 14 planes x 64 floats followed by the 64 label bytes of those pixels
// (255 = not in validPixels), 3648 contiguous bytes.  One tile = one cp.async.bulk (TMA bulk copy) into shared memory.
// Every consumer of a tile is an order-independent integer sum over its pixels, so the ORDER of the pixels inside a tile is
// free: it is chosen for the producer.  linearise_kernel's thread l of a 4-thread group holds the pixels 4l .. 4l+3 of a
// 16-pixel group and stores them as two float2 per plane; pixel 4l + j sits at position 2l + (j & 1) + 8 (j >> 1) of the group, so
// that four neighbouring threads fill one whole 32-byte sector with each store instruction instead of every other 8 bytes
// (half-written sectors doubled the L1 -> L2 store traffic and bounded the kernel).  A 16-pixel group keeps its place in the
// tile: the 32 positions a consumer's warp reads together are still 32 consecutive pixels of an image row (few label groups).
constexpr int ROW_TILE = 64;
constexpr int TILE_ROW_BYTES = ROW_TILE * NROWPL * 4;    // 3584
constexpr int TILE_BYTES = TILE_ROW_BYTES + ROW_TILE;    // 3648 (multiple of 16)
__host__ __device__ __forceinline__ size_t tiles_per_pair(size_t P0) { return (P0 + ROW_TILE - 1) / ROW_TILE; }
__host__ __device__ __forceinline__ int tile_pos(int p) {                // position of pixel p inside its tile
    const int r = p % ROW_TILE;
#ifdef SF_TILE_LINEAR
    return r;
#else
    return (r & ~15) + ((r >> 1) & 1) * 8 + (((r >> 2) & 3) << 1) + (r & 1);
#endif
}
__host__ __device__ __forceinline__ size_t tile_row_off(int k, int p) {   // byte offset of row plane k of pixel p
    return (size_t)(p / ROW_TILE) * TILE_BYTES + ((size_t)k * ROW_TILE + (size_t)tile_pos(p)) * 4;
}
__host__ __device__ __forceinline__ size_t tile_label_off(int p) {        // byte offset of the label of pixel p
    return (size_t)(p / ROW_TILE) * TILE_BYTES + TILE_ROW_BYTES + (size_t)tile_pos(p);
}

// fixed-point scales of the order-independent sums (mirrored by the oracle's EXACT policy)
constexpr int FIX_KMEANS = 36;  // k-means centre sums (KMeans.cpp:215)
constexpr int FIX_WARP_D = 32;  // warp depth accumulator (FrontEnd.cpp:840-867)
constexpr int FIX_WARP_I = 22;  // warp intensity accumulator, packed with the 22-bit weight sum
constexpr int FIX_PRIOR = 32;   // seg prior sum (SegmentationBackground.cpp:75)
constexpr int FIX_ABSB = 32;    // sum w|dt| for the initial mean residual (FrontEnd.cpp:590)
constexpr int FIX_RES = 30;     // per-label residual sums (FrontEnd.cpp:661)
// normal equations and |res|^2: columns scaled by powers of two below 2^QSCALE_BITS; the product of two scaled entries
// (exact in double) is rounded ONCE to a multiple of 2^-QFRAC_BITS and added, both inside one double-precision FMA onto an
// accumulator that stays in the binade of QMAGIC_D = 1.5 * 2^32: the ulp there is 2^-20, so fma(a, b, acc) ==
// acc + round_to_2^-20(a * b) exactly.  The accumulator's bit pattern minus QMAGIC_D's is the fixed-point sum as an integer;
// integer addition is associative -> any thread / block / GPU partition gives the same bits.  A thread may add at most
// 2^11 products (< 2^20 each) before its accumulator could leave the binade: MAX_TILES_PER_WARP_ITEM tiles (4 rows per lane).
constexpr int QSCALE_BITS = 10;
constexpr int QFRAC_BITS = 20;
constexpr double QMAGIC_D = 6442450944.0;                  // 1.5 * 2^32
constexpr long long QMAGIC_D_BITS = 0x41F8000000000000ll;  // its bit pattern
constexpr int MAX_TILES_PER_WARP_ITEM = 256;               // 1024 rows per thread: half the binade's head-room
constexpr int LABEL_BITS = 30;  // per-label residual terms: round(x * 2^s) < 2^30 (one cvt.rni.s32.f32), summed as two 16-bit limbs

// geometry of one pyramid level (host computes the float constants exactly as the reference does)
struct LevelGeom {
    int rows, cols, P;      // P = rows*cols
    float f;                // float(cols)/(2 tan(fovh/2))      FrontEnd.cpp:537,778
    float inv_f;            // 2 tan(fovh/2)/float(cols)        FrontEnd.cpp:378
    float inv_f_warp;       // 1.f/f                            FrontEnd.cpp:874
    float disp_u, disp_v;   // 0.5f*(cols-1), 0.5f*(rows-1)
    size_t off;             // offset (pixels) of this level inside a per-image pyramid
    unsigned cols_magic;    // ceil(2^32 / cols): pixel index -> (row, column) without an integer division
};

// row and column of pixel p of a level: p / cols == umulhi(p, ceil(2^32 / cols)) while p * cols < 2^32 (any level up to
// 2048 x 2048); the correction steps make the result right for every non-negative p anyway
__device__ __forceinline__ void split_rc(int p, const LevelGeom& g, int& row, int& col) {
    int q = (int)__umulhi((unsigned)p, g.cols_magic);
    int r = p - q * g.cols;
    if (r < 0) { q--; r += g.cols; }
    if (r >= g.cols) { q++; r -= g.cols; }
    row = q; col = r;
}

// solver tunables as the kernels see them
struct DevParams {
    int ctf_levels, max_iter_per_level, max_iter_irls, use_motion_filter, enable_segmentation;
    float k_photometric_res, irls_delta_threshold, kc_cauchy, kb, kz, lambda_reg, lambda_prior;
    float previous_speed_const_weight, previous_speed_eig_weight, outer_exit_threshold;
    float exp_neg_level[MAX_LEVELS];  // expf(-level) evaluated on the host (FrontEnd.cpp:745)
    float km_u_label[NC], km_v_label[NC];  // k-means seeds (KMeans.cpp:76-84), host-rounded
    float conn_dist2_threshold;       // square(0.03f*120.f/rows), KMeans.cpp:300
};

// per-pair control block, lives in global memory
struct PairCtl {
    // ---- pose ----
    float T[16];          // T_odometry, row-major
    float Tinv[12];       // rigid inverse, rows 0..2
    float twist_old[6];   // twist_odometry_old (input, then output after the frame)
    float twist_odom[6];  // twist_odometry = vee(log T_odometry)
    float twist_level[6]; // twist_level_odometry
    float var[6], prev_sol[6];
    double AtA[36];       // last weighted normal matrix (full, symmetric)
    double res_sq;        // ||res||^2 of the last IRLS iteration
    // ---- segmentation ----
    float b_segm[NC], b_prior[NC], lambda_t_w[NC];
    float kmeans[3 * NC];             // [3][24]: z, x, y
    float tbl_dist[NC * NC];          // sorted centre-to-centre distances (KMeans.cpp:249-259)
    unsigned char tbl_idx[NC * NC];
    unsigned conn[NC];                // adjacency bitmask rows (KMeans.cpp:308-340)
    // ---- per-step reductions ----
    unsigned max_wc_bits, max_wd_bits;
    long long fixBc, fixBd;
    long long prior_fix[NC];
    int csize[NC], cnonnull[NC];
    int n_valid;
    long long lab_fix[NC];
    int lab_cnt[NC];
    float inv_max_c, inv_max_d, aver_res, aver_res_old;
    unsigned colmax_c[7], colmax_d[7];  // max |row entry| per column (6 Jacobian columns + rhs) with raw weights, float bits
    float colbound[7];                  // bound of the normalised columns
    int sexp[7];                        // power-of-two column scales
    float mcs[7], mds[7];               // inv_max_c * 2^sexp[k], inv_max_d * 2^sexp[k] (exact scalings)
    int rexp;                           // power-of-two scale of the residuals
    long long acc_ne[27];               // integer normal equations: 21 upper-triangular AtA terms + 6 AtB terms
    long long acc_rs;                   // integer |res|^2
    // ---- control ----
    int active;        // this pair takes part in the current step
    int irls_done;     // IRLS loop of the current step has exited
    int break_level;   // level whose k-loop has been left (FrontEnd.cpp:1130), -1 = none
    int it_done;       // IRLS iterations done in the current step
    int status, total_irls;
    int km_iters;      // Lloyd iterations of the last kMeans3DCoord (KMeans.cpp:167-228)
    int items_cur;     // items the pair's current IRLS pass was cut into (irls_loop_kernel: ticket target)
    unsigned ticket1, ticket2;
    // ---- 5-frame history (computeResidualsAgainstPreviousImage, FrontEnd.cpp:896-1069) ----
    float Thist[12];          // rows 0..2 of (prod odomBuffer * T_odometry)^-1
    int hist_on;              // this pair has a 5-frame history in the current call
    int hist_ref;             // frame index of the image to warp (into the ring or the pyramid frames)
    long long hist_sum[NC];   // per-cluster sum of |dz| + k|dI|, 2^-32 fixed point
    int hist_cnt[NC];
};

// ---------------------------------------------------------------------------------------------
// fixed point helpers
// ---------------------------------------------------------------------------------------------
// round(x * 2^s): the multiplication by a power of two is exact (no overflow for the value ranges of the sums, gradual
// underflow is exact as well), so this equals llrint(ldexp(x, s)) of the oracle without ldexpf's special-case code
__device__ __forceinline__ long long fixq(float x, int s) { return __float2ll_rn(x * __int_as_float((127 + s) << 23)); }
// x * 2^e for |e| <= 1022 and no over/underflow of the result: identical to ldexp, without its slow path
__device__ __forceinline__ double scale_pow2(double x, int e) { return x * __hiloint2double((1023 + e) << 20, 0); }
__device__ __forceinline__ double fixval(long long acc, int s) { return scale_pow2((double)acc, -s); }
__host__ __device__ __forceinline__ int scale_exponent(float bound) {
    if (!(bound > 0.f) || !isfinite(bound)) return 0;
    int e;
    (void)frexpf(bound, &e);  // bound = m * 2^e, m in [0.5, 1)
    int s = QSCALE_BITS - e;
    if (s > 100) s = 100;
    if (s < -100) s = -100;
    return s;
}

// ---------------------------------------------------------------------------------------------
// small dense algebra in double — same operation sequences as the oracle's templates
// ---------------------------------------------------------------------------------------------
// (all loops fully unrolled: with N = 6 every index is a compile-time constant and the arrays live in registers)
template <int N>
__device__ __forceinline__ int ldlt_factor(double* A, unsigned char* zero) {
    const double tiny = 1e-20;
    int nz = 0;
#pragma unroll
    for (int j = 0; j < N; j++) {
        double d = A[j * N + j];
#pragma unroll
        for (int k = 0; k < j; k++) d -= (A[j * N + k] * A[j * N + k]) * A[k * N + k];
        A[j * N + j] = d;
        const bool z = !(d > tiny);
        zero[j] = z ? 1 : 0;
        nz += z ? 1 : 0;
#pragma unroll
        for (int i = j + 1; i < N; i++) {
            double s = A[i * N + j];
#pragma unroll
            for (int k = 0; k < j; k++) s -= (A[i * N + k] * A[j * N + k]) * A[k * N + k];
            A[i * N + j] = z ? 0.0 : s / d;
        }
    }
    return nz;
}
template <int N>
__device__ __forceinline__ void ldlt_solve_factored(const double* A, const unsigned char* zero, const double* b, double* x) {
#pragma unroll
    for (int i = 0; i < N; i++) {
        double s = b[i];
#pragma unroll
        for (int k = 0; k < i; k++) s -= A[i * N + k] * x[k];
        x[i] = s;
    }
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = zero[i] ? 0.0 : x[i] / A[i * N + i];
#pragma unroll
    for (int i = N - 1; i >= 0; i--) {  // descending k, as in the oracle
        double s = x[i];
#pragma unroll
        for (int k = N - 1; k > i; k--) s -= A[k * N + i] * x[k];
        x[i] = s;
    }
}

// Warp-cooperative LDL^T + solve of the 24x24 segmentation system held in shared memory
// (row stride LD).  Lane i owns row i; every element sees the same operation sequence as the
// serial ldlt_factor above, so the result is bit-identical to it.
template <int LD>
__device__ inline void ldlt24_warp(double* A, double* rhs, double* x, unsigned char* zero, int lane) {
    // Right-looking form: once column k is final, every remaining element (i, j > k) of the lower triangle gets its
    // k-th term subtracted -- independent updates instead of one dependent chain per element, and the same sequence
    // s -= (L[i][k] * L[j][k]) * D[k], k ascending, as the serial code sees.
    const double tiny = 1e-20;
    __syncwarp();
    for (int k = 0; k < NC; k++) {
        const double d = A[k * LD + k];  // final: all terms k' < k have been subtracted
        const bool z = !(d > tiny);
        if (lane == 0) zero[k] = z ? 1 : 0;
        double lik = 0.0;
        if (lane > k && lane < NC) {
            lik = z ? 0.0 : A[lane * LD + k] / d;
            A[lane * LD + k] = lik;
        }
        __syncwarp();
        if (lane > k && lane < NC) {
#pragma unroll 4
            for (int j = k + 1; j <= lane; j++) A[lane * LD + j] -= (lik * A[j * LD + k]) * d;
        }
        __syncwarp();
    }
    // triangular solves as column sweeps: lane i owns component i; every component sees the same sequence of
    // subtractions as the serial code (forward: k ascending, backward: k descending)
    double yi = (lane < NC) ? rhs[lane] : 0.0;
    for (int k = 0; k < NC; k++) {
        const double yk = __shfl_sync(0xffffffffu, yi, k);  // final: all its terms k' < k have been subtracted
        if (lane > k && lane < NC) yi -= A[lane * LD + k] * yk;
    }
    double xi = 0.0;
    if (lane < NC) xi = zero[lane] ? 0.0 : yi / A[lane * LD + lane];
    for (int k = NC - 1; k >= 0; k--) {
        const double xk = __shfl_sync(0xffffffffu, xi, k);
        if (lane < k) xi -= A[k * LD + lane] * xk;
    }
    if (lane < NC) x[lane] = xi;
    __syncwarp();
}

// cyclic Jacobi, symmetric 6x6 (stands in for SelfAdjointEigenSolver, FrontEnd.cpp:719)
__device__ inline void jacobi_eig6(const double* Ain, double* ev, double* V) {
    const int n = 6;
    double A[36];
    for (int i = 0; i < 36; i++) A[i] = Ain[i];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) V[i * n + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0, diag = 0;
        for (int p = 0; p < n; p++) {
            diag += A[p * n + p] * A[p * n + p];
            for (int q = p + 1; q < n; q++) off += A[p * n + q] * A[p * n + q];
        }
        if (!(off > 1e-32 * diag)) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = A[p * n + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0);
                const double s = t * c;
                for (int k = 0; k < n; k++) {
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    const double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; i++) ev[i] = A[i * n + i];
}

__device__ inline void mat3_sq(const double* K, double* K2) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += K[i * 3 + k] * K[k * 3 + j];
            K2[i * 3 + j] = s;
        }
}

// SE(3) exponential (stands in for Matrix4f::exp(), FrontEnd.cpp:766)
__device__ inline void se3_exp(const double* xi, double* M) {
    const double wx = xi[3], wy = xi[4], wz = xi[5];
    const double th2 = wx * wx + wy * wy + wz * wz;
    double A, B, C;
    if (th2 < 0.0025) {
        A = 1.0 + th2 * (-1.0 / 6 + th2 * (1.0 / 120 + th2 * (-1.0 / 5040 + th2 * (1.0 / 362880))));
        B = 0.5 + th2 * (-1.0 / 24 + th2 * (1.0 / 720 + th2 * (-1.0 / 40320 + th2 * (1.0 / 3628800))));
        C = 1.0 / 6 + th2 * (-1.0 / 120 + th2 * (1.0 / 5040 + th2 * (-1.0 / 362880 + th2 * (1.0 / 39916800))));
    } else {
        const double th = sqrt(th2);
        const double sh = sin(0.5 * th);
        A = sin(th) / th;
        B = 2.0 * sh * sh / th2;
        C = (th - sin(th)) / (th2 * th);
    }
    const double K[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double K2[9];
    mat3_sq(K, K2);
    double R[9], Vm[9];
    for (int i = 0; i < 9; i++) {
        const double I = (i % 4 == 0) ? 1.0 : 0.0;
        R[i] = I + A * K[i] + B * K2[i];
        Vm[i] = I + B * K[i] + C * K2[i];
    }
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) M[i * 4 + j] = R[i * 3 + j];
        M[i * 4 + 3] = Vm[i * 3 + 0] * xi[0] + Vm[i * 3 + 1] * xi[1] + Vm[i * 3 + 2] * xi[2];
    }
    M[12] = 0; M[13] = 0; M[14] = 0; M[15] = 1;
}

// SE(3) logarithm (stands in for Matrix4f::log(), FrontEnd.cpp:736,769)
__device__ inline void se3_log(const double* M, double* xi) {
    const double sx = 0.5 * (M[2 * 4 + 1] - M[1 * 4 + 2]);
    const double sy = 0.5 * (M[0 * 4 + 2] - M[2 * 4 + 0]);
    const double sz = 0.5 * (M[1 * 4 + 0] - M[0 * 4 + 1]);
    const double s2 = sx * sx + sy * sy + sz * sz;
    const double c = 0.5 * (M[0] + M[5] + M[10] - 1.0);
    double fac, th2, D;
    if (s2 < 0.0025 && c > 0.0) {
        fac = 1.0 + s2 * (1.0 / 6 + s2 * (3.0 / 40 + s2 * (15.0 / 336 + s2 * (105.0 / 3456 + s2 * (945.0 / 42240)))));
        th2 = s2 * fac * fac;
        D = 1.0 / 12 + th2 * (1.0 / 720 + th2 * (1.0 / 30240 + th2 * (1.0 / 1209600)));
    } else {
        const double s = sqrt(s2);
        const double th = atan2(s, c);
        fac = (s > 0.0) ? th / s : 1.0;
        th2 = th * th;
        const double sh = sin(0.5 * th);
        const double Bc = 2.0 * sh * sh / th2;
        const double Ac = sin(th) / th;
        D = (1.0 - Ac / (2.0 * Bc)) / th2;
    }
    const double wx = fac * sx, wy = fac * sy, wz = fac * sz;
    const double K[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double K2[9];
    mat3_sq(K, K2);
    for (int i = 0; i < 3; i++) {
        double acc = 0;
        for (int j = 0; j < 3; j++) {
            const double I = (i == j) ? 1.0 : 0.0;
            acc += (I - 0.5 * K[i * 3 + j] + D * K2[i * 3 + j]) * M[j * 4 + 3];
        }
        xi[i] = acc;
    }
    xi[3] = wx; xi[4] = wy; xi[5] = wz;
}

// rigid inverse of the float pose, rows 0..2 (stands in for Matrix4f::inverse(), FrontEnd.cpp:800)
__device__ inline void rigid_inverse(const float* M, float* out12) {
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) out12[i * 4 + j] = M[j * 4 + i];
        double s = 0;
        for (int j = 0; j < 3; j++) s += (double)M[j * 4 + i] * (double)M[j * 4 + 3];
        out12[i * 4 + 3] = (float)(-s);
    }
}

}  // namespace sf
