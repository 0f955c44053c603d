// sf_k_io.cu - the rows either side of the solver: depth pre-filter (getFilteredDepth) and image-sequence conversion (loadImageFromSequenceAssoc)
// Part of the sm_100a kernels of the StaticFusion joint odometry + segmentation solver (launch interface: sf_kernels.cuh).
// One launch of each kernel serves the whole batch of frame pairs; data-dependent exits (IRLS convergence FrontEnd.cpp:679,
// outer-loop exit :1130, k-means :227) are per-pair flags in PairCtl that later launches test, so the host enqueues a static
// schedule with no synchronisation.  Compiled with -fmad=false: float expressions keep the reference's operation order and
// rounding; fused multiply-adds appear only where written explicitly.  Reference citations are relative to the upstream tree.
#include <cstdlib>

#include "sf_common.cuh"

namespace sf {

// ------------------------------------------------------------------------------------------
// K0: depth pre-filter (Shaders/depth_bilateral.frag:30-76 + depth_metric.frag:28-40 as chained by
// Reconstruction::getFilteredDepth, Reconstruction.cpp:722-732): 13x13 bilateral on raw u16 millimetres.
// ------------------------------------------------------------------------------------------
// exp from IEEE float operations only == the oracle's det_expf bit for bit (nvcc -fmad=false keeps mul/add apart)
__device__ __forceinline__ float det_expf(float a) {
    if (!(a > -87.f)) return 0.f;
    if (a > 0.f) a = 0.f;
    const float k = rintf(a * 1.44269504088896341f);
    float r = a - k * 0.693359375f;
    r = r - k * -2.12194440e-4f;
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

constexpr int BF_R = 6, BF_TX = 32, BF_TY = 8;
__global__ void __launch_bounds__(BF_TX * BF_TY) filter_depth_kernel(const uint16_t* __restrict__ in, float* __restrict__ out, int rows, int cols,
                                                                     size_t in_stride, size_t out_stride, unsigned lim) {
    __shared__ uint16_t tile[BF_TY + 2 * BF_R][BF_TX + 2 * BF_R];
    const uint16_t* src = in + (size_t)blockIdx.z * in_stride;
    const int x0 = blockIdx.x * BF_TX - BF_R, y0 = blockIdx.y * BF_TY - BF_R;
    for (int i = threadIdx.y * BF_TX + threadIdx.x; i < (BF_TY + 2 * BF_R) * (BF_TX + 2 * BF_R); i += BF_TX * BF_TY) {
        const int ty = i / (BF_TX + 2 * BF_R), tx = i - ty * (BF_TX + 2 * BF_R);
        const int gx = x0 + tx, gy = y0 + ty;
        tile[ty][tx] = (gx >= 0 && gx < cols && gy >= 0 && gy < rows) ? src[(size_t)gy * cols + gx] : (uint16_t)0;
    }
    __syncthreads();
    const int x = blockIdx.x * BF_TX + threadIdx.x, y = blockIdx.y * BF_TY + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const unsigned value = tile[threadIdx.y + BF_R][threadIdx.x + BF_R];
    unsigned filtered = 0;
    if (!(value > lim || value < 300u)) {
        const float sigma_space2_inv_half = 0.024691358f;
        const float sigma_color2_inv_half = 0.000555556f;
        const int D = BF_R * 2 + 1;
        const int tx = min(x - D / 2 + D, cols), ty = min(y - D / 2 + D, rows);
        float sum1 = 0.f, sum2 = 0.f;
        for (int cy = max(y - D / 2, 0); cy < ty; ++cy)
            for (int cx = max(x - D / 2, 0); cx < tx; ++cx) {
                const unsigned tmp = tile[cy - y0][cx - x0];
                const float space2 = (float(x) - float(cx)) * (float(x) - float(cx)) + (float(y) - float(cy)) * (float(y) - float(cy));
                const float color2 = (float(value) - float(tmp)) * (float(value) - float(tmp));
                const float weight = det_expf(-(space2 * sigma_space2_inv_half + color2 * sigma_color2_inv_half));
                sum1 += float(tmp) * weight;
                sum2 += weight;
            }
        filtered = (unsigned)roundf(sum1 / sum2);
    }
    out[(size_t)blockIdx.z * out_stride + (size_t)y * cols + x] = (filtered > lim || filtered < 300u) ? 0.f : float(filtered) / 1000.0f;
}

// ------------------------------------------------------------------------------------------
// L1: image-sequence loader, conversion half (StaticFusion::loadImageFromSequenceAssoc, FrontEnd.cpp:216-254).
// Decoded 8-bit BGR + 16-bit depth at full resolution -> intensity [0,1], depth in metres, depth_mm and the
// colour image, vertically flipped (row H*rf - rf*v - 1) and decimated by res_factor (:231,249).  One thread per output
// pixel; each output is optional.  HBM-bound: 5 source bytes read (one 32 B sector per pixel when rf >= 2), up to 13 written.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) convert_frames_kernel(const uint8_t* __restrict__ bgr, const uint16_t* __restrict__ depth_raw, int rows,
                                                             int cols, int rf, float* __restrict__ intensity, float* __restrict__ depth,
                                                             size_t f32_stride, uint16_t* __restrict__ depth_mm, uint8_t* __restrict__ color) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= cols) return;
    const size_t full_cols = (size_t)cols * rf, P = (size_t)rows * cols, Pf = P * rf * rf;
    const size_t src = (size_t)blockIdx.z * Pf + (size_t)(rows * rf - rf * v - 1) * full_cols + (size_t)rf * u;
    const size_t dst = (size_t)v * cols + u;
    const float norm_factor = 1.f / 255.f;
    if (bgr) {
        const float r = norm_factor * (float)bgr[3 * src + 0];  // :232-234: channel 0 of cv::imread's BGR is named r
        const float g = norm_factor * (float)bgr[3 * src + 1];
        const float b = norm_factor * (float)bgr[3 * src + 2];
        if (intensity) intensity[(size_t)blockIdx.z * f32_stride + dst] = 0.299f * r + 0.587f * g + 0.114f * b;  // :236
        if (color) {  // :237, float -> uchar truncation
            uint8_t* o = color + 3 * ((size_t)blockIdx.z * P + dst);
            o[0] = (uint8_t)(r * 255); o[1] = (uint8_t)(g * 255); o[2] = (uint8_t)(b * 255);
        }
    }
    if (depth_raw) {
        const uint16_t raw = depth_raw[src];
        if (depth) depth[(size_t)blockIdx.z * f32_stride + dst] = (float)raw * 0.001f;  // :243,249 convertTo(CV_32FC1, 1/1000)
        if (depth_mm) depth_mm[(size_t)blockIdx.z * P + dst] = raw;                      // :244,250
    }
}

int launch_filter_depth(const uint16_t* in, float* out, int rows, int cols, int n, size_t in_stride, size_t out_stride, float max_depth_m,
                        cudaStream_t stream) {
    const unsigned lim = (unsigned)(max_depth_m * 1000.0f);
    filter_depth_kernel<<<dim3(cdiv(cols, BF_TX), cdiv(rows, BF_TY), n), dim3(BF_TX, BF_TY), 0, stream>>>(in, out, rows, cols, in_stride, out_stride, lim);
    return 1;
}

int launch_convert_frames(const uint8_t* bgr, const uint16_t* depth_raw, int rows, int cols, int res_factor, int n, float* intensity, float* depth,
                          size_t f32_stride, uint16_t* depth_mm, uint8_t* color, cudaStream_t stream) {
    convert_frames_kernel<<<dim3(cdiv(cols, 256), rows, n), 256, 0, stream>>>(bgr, depth_raw, rows, cols, res_factor, intensity, depth, f32_stride,
                                                                             depth_mm, color);
    return 1;
}
}  // namespace sf
