// sf_k_linearise.cu - linearisation: calculateCoord + calculateDerivatives + computeWeights + computeSegPrior sums + the Jacobian rows
// Part of the sm_100a kernels of the StaticFusion joint odometry + segmentation solver (launch interface: sf_kernels.cuh).
// One launch of each kernel serves the whole batch of frame pairs; data-dependent exits (IRLS convergence FrontEnd.cpp:679,
// outer-loop exit :1130, k-means :227) are per-pair flags in PairCtl that later launches test, so the host enqueues a static
// schedule with no synchronisation.  Compiled with -fmad=false: float expressions keep the reference's operation order and
// rounding; fused multiply-adds appear only where written explicitly.  Reference citations are relative to the upstream tree.
#include <cstdlib>

#include "sf_common.cuh"

namespace sf {

// ------------------------------------------------------------------------------------------
// Jacobian rows.  The 2N x 6 Jacobian A, B, Aw, Bw, res of the reference (FrontEnd.cpp:525-586) are never
// stored: both IRLS passes rebuild the two rows of a pixel in registers from the 11 linearisation scalars.
// ------------------------------------------------------------------------------------------
struct Rows {
    float ac[6], bc, ad[6], bd;
};

__device__ __forceinline__ void build_rows(float d, float x, float y, float dcu, float dcv, float dct, float ddu, float ddv,
                                           float ddt, float wc_raw, float wd_raw, float inv_max_c, float inv_max_d,
                                           float k_photo, float f_inv, Rows& r) {
    const float inv_d = rcp_rn_normal(d);  // d is a valid depth: normal range
    const float wc_n = inv_max_c * wc_raw;  // FrontEnd.cpp:505-509
    const float wd_n = inv_max_d * wd_raw;
    // colour, :552-565
    const float dycomp_c = dcu * f_inv * inv_d;
    const float dzcomp_c = dcv * f_inv * inv_d;
    const float twc = wc_n * k_photo;
    r.ac[0] = twc * (-dycomp_c);
    r.ac[1] = twc * (-dzcomp_c);
    r.ac[2] = twc * (dycomp_c * x * inv_d + dzcomp_c * y * inv_d);
    r.ac[3] = twc * (dycomp_c * inv_d * y * x + dzcomp_c * (y * y * inv_d + d));
    r.ac[4] = twc * (-dycomp_c * (x * x * inv_d + d) - dzcomp_c * inv_d * y * x);
    r.ac[5] = twc * (dycomp_c * y - dzcomp_c * x);
    r.bc = twc * (-dct);
    // geometry, :570-584
    const float dycomp_d = ddu * f_inv * inv_d;
    const float dzcomp_d = ddv * f_inv * inv_d;
    const float twd = wd_n;
    r.ad[0] = twd * (-dycomp_d);
    r.ad[1] = twd * (-dzcomp_d);
    r.ad[2] = twd * (1.f + dycomp_d * x * inv_d + dzcomp_d * y * inv_d);
    r.ad[3] = twd * (y + dycomp_d * inv_d * y * x + dzcomp_d * (y * y * inv_d + d));
    r.ad[4] = twd * (-x - dycomp_d * (x * x * inv_d + d) - dzcomp_d * inv_d * y * x);
    r.ad[5] = twd * (dycomp_d * y - dzcomp_d * x);
    r.bd = twd * (-ddt);
}

// ------------------------------------------------------------------------------------------
// K3: linearisation = calculateCoord + calculateDerivatives + computeWeights (raw) +
// computeSegPrior sums (FrontEnd.cpp:393-510, SegmentationBackground.cpp:53-81)
// ------------------------------------------------------------------------------------------
// 4 horizontally adjacent pixels per thread (float4 loads / stores); the per-pixel expressions are literal.  All
// reductions are integer sums or maxima, accumulated in registers over the 4 pixels, then per warp, per block, per pair.
#ifndef SF_LIN_THREADS
#define SF_LIN_THREADS 256
#endif
#ifndef SF_LIN_BPS
#define SF_LIN_BPS 2  // resident blocks per SM: 128 registers, no spills (3 blocks = 85 registers spills ~400 B per thread and is 1.5x slower)
#endif
// NS > 0 (staged): the four source planes of an item (its pixels plus one image row above and below: one contiguous range per
// plane) arrive in shared memory by four bulk copies (cp.async.bulk, mbarrier completion) issued NS - 1 items ahead into a ring
// of NS stages, so DRAM latency is hidden by the copy engine instead of by resident warps (the kernel needs 128 registers = 16
// warps per SM).  A stage is handed back by its own "empty" mbarrier (one arrival per warp): only the issuing thread waits for
// it, the warps of a block drift up to NS - 1 items apart and there is no block-wide barrier per item.  NS = 0 (read-only
// global loads) serves the levels whose stages would not fit.
constexpr int LIN_ITEM_PIXELS = SF_LIN_THREADS * 4;
constexpr int LIN_WARPS = SF_LIN_THREADS / 32;
constexpr int LIN_MAX_STAGES = 3;
__host__ __device__ __forceinline__ int lin_span(int cols) { return LIN_ITEM_PIXELS + 2 * cols; }  // floats of one plane of one stage
__host__ __device__ __forceinline__ size_t lin_dyn_smem(int cols, int ns) { return (size_t)ns * 4 * lin_span(cols) * sizeof(float) + 2 * LIN_MAX_STAGES * sizeof(unsigned long long); }
template <int NS>
__global__ void __launch_bounds__(SF_LIN_THREADS, SF_LIN_BPS) linearise_kernel(Arena a, DevParams prm, LevelGeom g, int first, int blocks_per_pair) {
    constexpr bool STAGED = NS > 0;
    // persistent grid over (active pair, 1024-pixel block) items.  A block takes a CONTIGUOUS range of items, i.e. mostly one
    // pair: the per-thread partial reductions stay in registers across items and are folded (warp -> block -> PairCtl) only
    // when the pair changes, not once per 4 pixels of every thread.  All sums are integer sums / maxima: any cut is exact.
    const int total = a.gcount[0] * blocks_per_pair;
    const int tid = threadIdx.x, lane = tid & 31;
    __shared__ long long s_prior[NC];
    // in-loop flushes of a finished label run add the 64-bit prior term as three limbs (bits 0-15, 16-31, 32+) with native
    // 32-bit shared atomics: a 64-bit shared atomic is a CAS spin loop and was the kernel's top stall.  At most 128 items
    // x 256 threads flushes of < 2^16 between two folds: no limb overflows
    __shared__ unsigned s_prior_lo[NC], s_prior_mid[NC];
    __shared__ int s_prior_hi[NC];
    __shared__ int s_size[NC], s_nonnull[NC];
    __shared__ long long s_fixBc, s_fixBd;
    __shared__ unsigned s_maxc, s_maxd;
    __shared__ unsigned s_colmax[14];
    __shared__ int s_nvalid;
    const int item0 = (int)(((long long)blockIdx.x * total) / gridDim.x), item1 = (int)(((long long)(blockIdx.x + 1) * total) / gridDim.x);
    extern __shared__ __align__(128) unsigned char lin_smem[];
    const int span = lin_span(g.cols);
    float* const stage_mem = reinterpret_cast<float*>(lin_smem);
    unsigned long long* const full_bar = reinterpret_cast<unsigned long long*>(lin_smem + (size_t)NS * 4 * span * sizeof(float));
    unsigned long long* const empty_bar = full_bar + LIN_MAX_STAGES;
    // first pixel / number of pixels an item needs of every plane, and the copies themselves (thread 0)
    auto item_range = [&](int it_, int& lo, int& n) {
        const int ip0 = (it_ - (it_ / blocks_per_pair) * blocks_per_pair) * LIN_ITEM_PIXELS;
        lo = max(0, ip0 - g.cols);
        n = min(g.P, ip0 + LIN_ITEM_PIXELS + g.cols) - lo;
    };
    // the issuing thread keeps (slot, pair, current frame, predicted frame) of the last issued item in shared memory: the two
    // dependent global loads behind them would otherwise hold warp 0 - and, one ring later, every warp - for ~2 us per item
    __shared__ int s_iss[4];
    if (tid == 0) s_iss[0] = -1;
    auto issue_item = [&](int it_, int st) {
        const int slot_ = it_ / blocks_per_pair;
        if (slot_ != s_iss[0]) {
            const int pr_ = a.active_list[slot_];
            s_iss[0] = slot_; s_iss[1] = pr_; s_iss[2] = a.cur_idx[pr_]; s_iss[3] = a.pred_idx[pr_];
        }
        const int pr = s_iss[1], fc_ = s_iss[2], fp_ = s_iss[3];
        int lo, n;
        item_range(it_, lo, n);
        const float* src[4] = {a.pyr_d + (size_t)fc_ * a.pyr_stride + g.off, a.pyr_i + (size_t)fc_ * a.pyr_stride + g.off,
                               first ? a.pyr_d + (size_t)fp_ * a.pyr_stride + g.off : a.warp_d + (size_t)pr * a.P0,
                               first ? a.pyr_i + (size_t)fp_ * a.pyr_stride + g.off : a.warp_i + (size_t)pr * a.P0};
        const unsigned bytes = (unsigned)n * 4u;  // lo and n are multiples of 4 pixels: 16-byte aligned, 16-byte granular
        mbar_arm(&full_bar[st], 4u * bytes);
#pragma unroll
        for (int q = 0; q < 4; q++) bulk_load(stage_mem + ((size_t)st * 4 + q) * span, src[q] + lo, bytes, &full_bar[st]);
    };
    if (STAGED) {
        if (tid < NS) { mbar_init(&full_bar[tid], 1); mbar_init(&empty_bar[tid], LIN_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
        if (tid == 0)
            for (int q = 0; q < NS - 1; q++)
                if (item0 + q < item1) issue_item(item0 + q, q);
    }
    // ring state.  Consumers: stage / parity of the current item.  Issuing thread: the stage the item NS - 1 ahead goes into (the one
    // read at the previous item) and the parity of that stage's last hand-back
    int st = 0;
    unsigned st_par = 0;
    int iss_st = NS - 1;
    unsigned iss_par = 1;

    // per-thread partial reductions of the current pair
    float t_maxc = 0.f, t_maxd = 0.f;
    float t_colmax[14];
#pragma unroll
    for (int q = 0; q < 14; q++) t_colmax[q] = 0.f;
    long long t_qBc = 0, t_qBd = 0;
    int t_nvalid = 0;
    int run_lab = -1, run_size = 0, run_nonnull = 0;
    long long run_prior = 0;
    int cur_pair = -1, items_since_fold = 0;
    // (slot, rem) = the item's pair slot and 1024-pixel block in it; the pair of the slot stays in a register (one dependent
    // global load per pair, not per item) and the labels of an item are loaded one item ahead (their DRAM latency was the
    // kernel's top stall: nothing else of an item comes from global memory)
    int slot = item0 / blocks_per_pair, rem = item0 - slot * blocks_per_pair;
    int pair_next = item0 < item1 ? a.active_list[slot] : 0;
    // the prefetched labels stay ONE packed 32-bit register until the item that uses them: unpacked where they are loaded (as a
    // uchar4 is), the byte extraction waits for the load at once
    const unsigned none4 = 0x01010101u * LABEL_NONE;
    auto load_labels = [&](int pr, int rem_) -> unsigned {
        const int ch = rem_ * SF_LIN_THREADS + tid;
        if (ch < (g.P >> 2)) return __ldg(reinterpret_cast<const unsigned*>(a.labels + (size_t)pr * a.pyr_stride + g.off + ((size_t)ch << 2)));
        return none4;
    };
    unsigned l4n = item0 < item1 ? load_labels(pair_next, rem) : none4;

    // fold the block's partial reductions into the pair's cells (block-collective) and reset them
    auto flush = [&](int pair) {
        PairCtl& c = a.ctl[pair];
        warp_bins_add(run_size ? run_lab : -1, run_prior, 0, 0, run_nonnull, s_prior, nullptr, nullptr, nullptr, s_nonnull, lane);
        {   // sizes of the last runs (warp_bins_add's first counter counts lanes, not pixels)
            unsigned todo = __ballot_sync(0xffffffffu, run_size > 0);
            while (todo) {
                const int leader = __ffs(todo) - 1;
                const int l = __shfl_sync(0xffffffffu, run_lab, leader);
                const bool mine = (run_size > 0) && (run_lab == l);
                const unsigned grp = __ballot_sync(0xffffffffu, mine);
                const int n = __reduce_add_sync(0xffffffffu, mine ? run_size : 0);
                if (lane == leader) atomicAdd(&s_size[l], n);
                todo &= ~grp;
            }
        }
        const int nv = __reduce_add_sync(0xffffffffu, t_nvalid);
        if (nv) {  // warp-uniform
            const unsigned mc = __reduce_max_sync(0xffffffffu, __float_as_uint(t_maxc));
            const unsigned md = __reduce_max_sync(0xffffffffu, __float_as_uint(t_maxd));
            unsigned cm[14];
#pragma unroll
            for (int q = 0; q < 14; q++) cm[q] = __reduce_max_sync(0xffffffffu, __float_as_uint(t_colmax[q]));
            const long long sBc = warp_sum_ll(t_qBc), sBd = warp_sum_ll(t_qBd);
            if (lane == 0) {
                atomicMax(&s_maxc, mc); atomicMax(&s_maxd, md);
                atomic_add_ll(&s_fixBc, sBc); atomic_add_ll(&s_fixBd, sBd);
                atomicAdd(&s_nvalid, nv);
#pragma unroll
                for (int q = 0; q < 14; q++) atomicMax(&s_colmax[q], cm[q]);
            }
        }
        __syncthreads();
        if (tid < NC) {
            if (s_size[tid]) atomicAdd(&c.csize[tid], s_size[tid]);
            if (s_nonnull[tid]) atomicAdd(&c.cnonnull[tid], s_nonnull[tid]);
            const long long pr = s_prior[tid] + (long long)s_prior_lo[tid] + ((long long)s_prior_mid[tid] << 16) + ((long long)s_prior_hi[tid] << 32);
            if (pr) atomic_add_ll(&c.prior_fix[tid], pr);
        }
        if (tid < 14 && s_nvalid) {
            if (tid < 7) atomicMax(&c.colmax_c[tid], s_colmax[tid]);
            else atomicMax(&c.colmax_d[tid - 7], s_colmax[tid]);
        }
        if (tid == 0 && s_nvalid) {
            atomicMax(&c.max_wc_bits, s_maxc); atomicMax(&c.max_wd_bits, s_maxd);
            atomic_add_ll(&c.fixBc, s_fixBc); atomic_add_ll(&c.fixBd, s_fixBd);
            atomicAdd(&c.n_valid, s_nvalid);
        }
        t_maxc = 0.f; t_maxd = 0.f;
#pragma unroll
        for (int q = 0; q < 14; q++) t_colmax[q] = 0.f;
        t_qBc = 0; t_qBd = 0; t_nvalid = 0;
        run_lab = -1; run_size = 0; run_nonnull = 0; run_prior = 0;
        __syncthreads();  // the block totals have been read: they may be reset
    };

  for (int item = item0; item < item1; item++) {
    const int pair = pair_next;
    const int chunk = rem * SF_LIN_THREADS + tid;
    const bool inb = chunk < (g.P >> 2);
    const unsigned l4 = l4n;
    if (++rem == blocks_per_pair) { rem = 0; slot++; if (item + 1 < item1) pair_next = a.active_list[slot]; }
    if (item + 1 < item1) l4n = load_labels(pair_next, rem);  // consumed at the next item
    if (pair != cur_pair || items_since_fold == 128) {  // block-uniform
        if (cur_pair >= 0) flush(cur_pair);
        cur_pair = pair;
        items_since_fold = 0;
        if (tid < NC) { s_prior[tid] = 0; s_prior_lo[tid] = 0; s_prior_mid[tid] = 0; s_prior_hi[tid] = 0; s_size[tid] = 0; s_nonnull[tid] = 0; }
        if (tid < 14) s_colmax[tid] = 0;
        if (tid == 0) { s_fixBc = 0; s_fixBd = 0; s_maxc = 0; s_maxd = 0; s_nvalid = 0; }
        __syncthreads();
    }

    int st_lo = 0;
    if (STAGED) {
        if (tid == 0) {
            if (item + (NS - 1) < item1) {
                if (item > item0) mbar_wait(&empty_bar[iss_st], iss_par);  // every warp is done with the previous item's stage
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue_item(item + (NS - 1), iss_st);
            }
            if (++iss_st == NS) { iss_st = 0; iss_par ^= 1u; }
        }
        int n_unused;
        item_range(item, st_lo, n_unused);
        mbar_wait(&full_bar[st], st_par);
    }
    const float* const sp = stage_mem + (size_t)st * 4 * span - st_lo;  // plane q of this item: sp[q * span + pixel]
    items_since_fold++;
    {   // pad pixels of the level's last, partial tile: stale labels of a finer level must not be read as valid
        const int padded = (int)tiles_per_pair((size_t)g.P) * (ROW_TILE / 4);
        if (!inb && chunk < padded) {
            uint8_t* tb = a.tiles + (size_t)pair * tiles_per_pair(a.P0) * TILE_BYTES;
            for (int h = 0; h < 4; h += 2) {  // pixels (0, 1) and (2, 3) of the chunk are neighbours inside the tile
                *reinterpret_cast<uchar2*>(tb + tile_label_off((chunk << 2) + h)) = make_uchar2(VLABEL_INVALID, VLABEL_INVALID);
                for (int k = 0; k < NROWPL; k++)  // the passes multiply invalid pixels by a zero weight: rows must be finite
                    *reinterpret_cast<float2*>(tb + tile_row_off(k, (chunk << 2) + h)) = make_float2(0.f, 0.f);
            }
        }
    }
    const int fc = a.cur_idx[pair], fp = a.pred_idx[pair];
    const float* cd = a.pyr_d + (size_t)fc * a.pyr_stride + g.off;
    const float* ci = a.pyr_i + (size_t)fc * a.pyr_stride + g.off;
    // the very first step uses the prediction level itself as the warped image (FrontEnd.cpp:1103-1110)
    const float* wdp = first ? a.pyr_d + (size_t)fp * a.pyr_stride + g.off : a.warp_d + (size_t)pair * a.P0;
    const float* wip = first ? a.pyr_i + (size_t)fp * a.pyr_stride + g.off : a.warp_i + (size_t)pair * a.P0;
    uint8_t* tiles = a.tiles + (size_t)pair * tiles_per_pair(a.P0) * TILE_BYTES;
    float* dbg = a.dbg ? a.dbg + (size_t)pair * NPLANES * a.P0 : nullptr;
    // plane q (0 depth, 1 intensity of the current frame; 2, 3 of the warped one) at pixel px: staged copy or global memory
    const float* const gsrc[4] = {cd, ci, wdp, wip};
    auto L4 = [&](int q, int px) -> float4 {
        if (STAGED) return *reinterpret_cast<const float4*>(sp + (size_t)q * span + px);
        return ldg4(gsrc[q] + px);
    };
    auto L1 = [&](int q, int px) -> float {
        if (STAGED) return sp[(size_t)q * span + px];
        return __ldg(gsrc[q] + px);
    };

    if (inb) {
        const int p0 = chunk << 2;
        int v, u0;
        split_rc(p0, g, v, u0);
        const bool has_up = v > 0, has_dn = v < g.rows - 1, has_l = u0 > 0, has_r = u0 + 4 < g.cols;
        // centre row: positions -1 .. 4
        float dcur[6], icur[6], dwar[6], iwar[6];
        {
            const float4 a0 = L4(0, p0), a1 = L4(1, p0), a2 = L4(2, p0), a3 = L4(3, p0);
            dcur[1] = a0.x; dcur[2] = a0.y; dcur[3] = a0.z; dcur[4] = a0.w;
            icur[1] = a1.x; icur[2] = a1.y; icur[3] = a1.z; icur[4] = a1.w;
            dwar[1] = a2.x; dwar[2] = a2.y; dwar[3] = a2.z; dwar[4] = a2.w;
            iwar[1] = a3.x; iwar[2] = a3.y; iwar[3] = a3.z; iwar[4] = a3.w;
            dcur[0] = has_l ? L1(0, p0 - 1) : 0.f; icur[0] = has_l ? L1(1, p0 - 1) : 0.f;
            dwar[0] = has_l ? L1(2, p0 - 1) : 0.f; iwar[0] = has_l ? L1(3, p0 - 1) : 0.f;
            dcur[5] = has_r ? L1(0, p0 + 4) : 0.f; icur[5] = has_r ? L1(1, p0 + 4) : 0.f;
            dwar[5] = has_r ? L1(2, p0 + 4) : 0.f; iwar[5] = has_r ? L1(3, p0 + 4) : 0.f;
        }
        // intermediate depth / intensity of the row and its vertical neighbours (0 depth where Null, :411-428)
        float dI[6], II[6];
        bool nul[6];
#pragma unroll
        for (int k = 0; k < 6; k++) {
            nul[k] = !((dcur[k] != 0.f) && (dwar[k] != 0.f));
            dI[k] = nul[k] ? 0.f : 0.5f * (dcur[k] + dwar[k]);
            II[k] = 0.5f * (icur[k] + iwar[k]);
        }
        float dU[4], IU[4], dD[4], ID[4];
        bool nU[4], nD[4];
        {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 u0d = has_up ? L4(0, p0 - g.cols) : z, u1 = has_up ? L4(1, p0 - g.cols) : z;
            const float4 u2 = has_up ? L4(2, p0 - g.cols) : z, u3 = has_up ? L4(3, p0 - g.cols) : z;
            const float4 d0 = has_dn ? L4(0, p0 + g.cols) : z, d1 = has_dn ? L4(1, p0 + g.cols) : z;
            const float4 d2 = has_dn ? L4(2, p0 + g.cols) : z, d3 = has_dn ? L4(3, p0 + g.cols) : z;
            const float uc[4] = {u0d.x, u0d.y, u0d.z, u0d.w}, ui[4] = {u1.x, u1.y, u1.z, u1.w};
            const float uw[4] = {u2.x, u2.y, u2.z, u2.w}, uwi[4] = {u3.x, u3.y, u3.z, u3.w};
            const float dc4[4] = {d0.x, d0.y, d0.z, d0.w}, di4[4] = {d1.x, d1.y, d1.z, d1.w};
            const float dw4[4] = {d2.x, d2.y, d2.z, d2.w}, dwi4[4] = {d3.x, d3.y, d3.z, d3.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                nU[j] = !((uc[j] != 0.f) && (uw[j] != 0.f));
                dU[j] = nU[j] ? 0.f : 0.5f * (uc[j] + uw[j]);
                IU[j] = 0.5f * (ui[j] + uwi[j]);
                nD[j] = !((dc4[j] != 0.f) && (dw4[j] != 0.f));
                dD[j] = nD[j] ? 0.f : 0.5f * (dc4[j] + dw4[j]);
                ID[j] = 0.5f * (di4[j] + dwi4[j]);
            }
        }
        const int ll[4] = {(int)(l4 & 255u), (int)((l4 >> 8) & 255u), (int)((l4 >> 16) & 255u), (int)(l4 >> 24)};
        float ro[NROWPL][2];  // rows of a pixel pair, stored as float2 (8 B per lane: full sectors)
        unsigned char ovl[4];
        const float cv = float(v) - g.disp_v;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int u = u0 + j;
            const float dc = dcur[j + 1], ic = icur[j + 1], dw = dwar[j + 1], iw = iwar[j + 1];
            const bool isnull = nul[j + 1];
            const float dct = ic - iw, ddt = dc - dw;  // :477-478
            const int l = ll[j];
            if (l != LABEL_NONE) {  // SegmentationBackground.cpp:68-80
                if (l != run_lab) {
                    if (run_size) {
                        atomicAdd(&s_size[run_lab], run_size);
                        if (run_nonnull) {
                            atomicAdd(&s_nonnull[run_lab], run_nonnull);
                            atomicAdd(&s_prior_lo[run_lab], (unsigned)(run_prior & 0xffff));
                            atomicAdd(&s_prior_mid[run_lab], (unsigned)((run_prior >> 16) & 0xffff));
                            atomicAdd(&s_prior_hi[run_lab], (int)(run_prior >> 32));
                        }
                    }
                    run_lab = l; run_size = 0; run_nonnull = 0; run_prior = 0;
                }
                run_size++;
                if (!isnull) { run_nonnull++; run_prior += fixq(1.f - prm.kz * fabsf(ddt), FIX_PRIOR); }
            }
#pragma unroll
            for (int k = 0; k < NROWPL; k++) ro[k][j & 1] = 0.f;
            if (dbg) { dbg[(size_t)PL_DCT * a.P0 + p0 + j] = dct; dbg[(size_t)PL_DDT * a.P0 + p0 + j] = ddt; }
            const bool valid = !isnull && (u != 0) && (v != 0) && (u != g.cols - 1) && (v != g.rows - 1);  // :417
            unsigned char vl = VLABEL_INVALID;
            if (valid) {
                const float cu = float(u) - g.disp_u;
                const float xc = (g.inv_f * cu) * dc, yc = (g.inv_f * cv) * dc;  // xxPyr / yyPyr
                float xw, yw;
                if (first) { xw = (g.inv_f * cu) * dw; yw = (g.inv_f * cv) * dw; }  // xxPredPyr
                else { xw = cu * dw * g.inv_f_warp; yw = cv * dw * g.inv_f_warp; }  // :883-884
                const float d = dI[j + 1];  // 0.5f*(dc+dw), :413
                const float x = 0.5f * (xc + xw);
                const float y = 0.5f * (yc + yw);
                const float I = II[j + 1];  // :428
                const float epsilon_intensity = 1e-6f, epsilon_depth = 0.005f;  // :445-446
                // rx(v,u), rx(v,u-1), ry(v,u), ry(v-1,u)  (:448-462; 1 where the pixel itself is Null)
                const float rx_c = fabsf(dI[j + 2] - d) + epsilon_depth;
                const float rxI_c = fabsf(II[j + 2] - I) + epsilon_intensity;
                const float rx_l = nul[j] ? 1.f : fabsf(d - dI[j]) + epsilon_depth;
                const float rxI_l = nul[j] ? 1.f : fabsf(I - II[j]) + epsilon_intensity;
                const float ry_c = fabsf(dD[j] - d) + epsilon_depth;
                const float ryI_c = fabsf(ID[j] - I) + epsilon_intensity;
                const float ry_u = nU[j] ? 1.f : fabsf(d - dU[j]) + epsilon_depth;
                const float ryI_u = nU[j] ? 1.f : fabsf(I - IU[j]) + epsilon_intensity;
                // :470-473
                const float dcu = div_rn_pos(rxI_l * (II[j + 2] - I) + rxI_c * (I - II[j]), rxI_c + rxI_l);
                const float ddu = div_rn_pos(rx_l * (dI[j + 2] - d) + rx_c * (d - dI[j]), rx_c + rx_l);
                const float dcv = div_rn_pos(ryI_u * (ID[j] - I) + ryI_c * (I - IU[j]), ryI_c + ryI_u);
                const float ddv = div_rn_pos(ry_u * (dD[j] - d) + ry_c * (d - dU[j]), ry_c + ry_u);
                // computeWeights, :494-503 (normalisation by the global maxima is applied where the weights are read)
                const float error_l_c = 10.f * (fabsf(dct) + fabsf(dcu) + fabsf(dcv));
                const float error_l_d = 200.f * (fabsf(ddt) + fabsf(ddu) + fabsf(ddv));
                const float wc = sqrt_rn_normal(rcp_rn_normal(1.f + error_l_c));
                const float wd = sqrt_rn_normal(rcp_rn_normal(0.01f + error_l_d));
                t_maxc = fmaxf(t_maxc, wc); t_maxd = fmaxf(t_maxd, wd);
                t_qBc += fixq(wc * fabsf(dct), FIX_ABSB);
                t_qBd += fixq(wd * fabsf(ddt), FIX_ABSB);
                t_nvalid++;
                Rows rr;  // rows built with the raw pre-weights: their column maxima bound the normalised system
                build_rows(d, x, y, dcu, dcv, dct, ddu, ddv, ddt, wc, wd, 1.f, 1.f, prm.k_photometric_res, g.f, rr);
#pragma unroll
                for (int q = 0; q < 6; q++) {
                    t_colmax[q] = fmaxf(t_colmax[q], fabsf(rr.ac[q]));
                    t_colmax[7 + q] = fmaxf(t_colmax[7 + q], fabsf(rr.ad[q]));
                }
                t_colmax[6] = fmaxf(t_colmax[6], fabsf(rr.bc));
                t_colmax[13] = fmaxf(t_colmax[13], fabsf(rr.bd));
#pragma unroll
                for (int q = 0; q < 6; q++) { ro[RW_AC + q][j & 1] = rr.ac[q]; ro[RW_AD + q][j & 1] = rr.ad[q]; }
                ro[RW_BC][j & 1] = rr.bc; ro[RW_BD][j & 1] = rr.bd;
                if (dbg) {
                    const size_t q0 = (size_t)p0 + j;
                    dbg[(size_t)PL_D * a.P0 + q0] = d; dbg[(size_t)PL_X * a.P0 + q0] = x; dbg[(size_t)PL_Y * a.P0 + q0] = y;
                    dbg[(size_t)PL_DCU * a.P0 + q0] = dcu; dbg[(size_t)PL_DCV * a.P0 + q0] = dcv;
                    dbg[(size_t)PL_DDU * a.P0 + q0] = ddu; dbg[(size_t)PL_DDV * a.P0 + q0] = ddv;
                    dbg[(size_t)PL_WC * a.P0 + q0] = wc; dbg[(size_t)PL_WD * a.P0 + q0] = wd;
                }
                vl = prm.enable_segmentation ? (unsigned char)l : (unsigned char)0;
            }
            ovl[j] = vl;
            if (j & 1) {
#pragma unroll
                for (int k = 0; k < NROWPL; k++)
                    *reinterpret_cast<float2*>(tiles + tile_row_off(k, p0 + (j - 1))) = make_float2(ro[k][0], ro[k][1]);
            }
        }
        *reinterpret_cast<uchar2*>(tiles + tile_label_off(p0)) = make_uchar2(ovl[0], ovl[1]);
        *reinterpret_cast<uchar2*>(tiles + tile_label_off(p0 + 2)) = make_uchar2(ovl[2], ovl[3]);
    }
    if (STAGED) {  // this warp is done with the item's stage: hand it back
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[st]);
        if (++st == NS) { st = 0; st_par ^= 1u; }
    }
  }
    if (cur_pair >= 0) flush(cur_pair);
}

// finalise the step's reductions: seg prior (SegmentationBackground.cpp:84-102), weight maxima
// (FrontEnd.cpp:505-509), initial mean residual (:589-590), b_segm initialisation (:603-607)
__global__ void step_prep_kernel(Arena a, DevParams prm, int level_i, int k, int n_pairs) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= n_pairs) return;
    PairCtl& c = a.ctl[pair];
    if (!c.active) return;
    float* tr = a.trace ? a.trace + ((size_t)pair * a.trace_steps + (level_i * prm.max_iter_per_level + k)) * SF_TRACE_STEP : nullptr;
    for (int l = 0; l < NC; l++) {
        float bp = 0.f, ltw = 0.f;
        if (c.csize[l] != 0) {
            const float ratio = float(c.cnonnull[l]) / float(c.csize[l]);
            if (ratio < 0.1f) { ltw = 0.1f; bp = -1.f; }
            else {
                ltw = ratio;
                const float mean = (float)(fixval(c.prior_fix[l], FIX_PRIOR) / (double)c.cnonnull[l]);
                bp = fmaxf(-1.f, fminf(2.f, mean));
            }
        }
        c.b_prior[l] = bp; c.lambda_t_w[l] = ltw;
    }
    const int N = c.n_valid;
    const float maxc = __uint_as_float(c.max_wc_bits), maxd = __uint_as_float(c.max_wd_bits);
    for (int i = 0; i < 6; i++) { c.var[i] = 0.f; c.prev_sol[i] = 0.f; }
    c.it_done = 0;
    for (int l = 0; l < NC; l++) { c.lab_fix[l] = 0; c.lab_cnt[l] = 0; }
    for (int q = 0; q < 27; q++) c.acc_ne[q] = 0;
    c.acc_rs = 0; c.rexp = 0;
    bool degenerate = false;
    float aver = 0.f;
    if (N == 0 || !(maxc > 0.f) || !(maxd > 0.f)) {
        c.status |= SF_STATUS_NO_VALID_PIXELS; degenerate = true;
    } else {
        c.inv_max_c = 1.f / maxc; c.inv_max_d = 1.f / maxd;
        for (int q = 0; q < 7; q++) {  // power-of-two column scales of the integer normal equations
            const float cb = fmaxf(c.inv_max_c * __uint_as_float(c.colmax_c[q]), c.inv_max_d * __uint_as_float(c.colmax_d[q]));
            c.colbound[q] = cb;
            c.sexp[q] = scale_exponent(cb);
            c.mcs[q] = ldexpf(c.inv_max_c, c.sexp[q]);  // exact: the power of two commutes with every later rounding
            c.mds[q] = ldexpf(c.inv_max_d, c.sexp[q]);
        }
        const double sc = (double)c.inv_max_c * (double)prm.k_photometric_res;
        aver = (float)((sc * fixval(c.fixBc, FIX_ABSB) + (double)c.inv_max_d * fixval(c.fixBd, FIX_ABSB)) / (double)(2 * N));
        if (!(aver > 0.f) || !isfinite(aver)) { c.status |= SF_STATUS_ZERO_RESIDUAL; degenerate = true; }
    }
    c.aver_res = aver; c.aver_res_old = aver;
    if (!degenerate) {
        if (!prm.enable_segmentation) for (int l = 0; l < NC; l++) c.b_segm[l] = 1.f;
        else if (level_i == 0) for (int l = 0; l < NC; l++) c.b_segm[l] = c.b_prior[l];
    }
    c.irls_done = degenerate ? 2 : 0;  // 2 = degenerate step: pose_update leaves T untouched
    if (!degenerate) {  // the pair enters the IRLS loop: first iteration's work list
        a.iter_list0[atomicAdd(&a.gcount[2], 1)] = pair;
        atomicAdd(&a.gcount[1], 1);
    }
    if (tr) {
        tr[0] = 1.f; tr[1] = (float)level_i; tr[2] = (float)k; tr[3] = (float)N;
        tr[5] = maxc; tr[6] = maxd; tr[7] = aver;
        for (int l = 0; l < NC; l++) { tr[8 + l] = c.b_prior[l]; tr[32 + l] = c.lambda_t_w[l]; }
    }
}

constexpr size_t LIN_MAX_STAGED_SMEM_PER_SM = 200 * 1024;  // the resident blocks' stages must fit beside their static shared memory
int launch_linearise(const Arena& a, const DevParams& p, const LevelGeom& g, int first, const LaunchCfg& c) {
    const int bpp = (int)cdiv(tiles_per_pair((size_t)g.P) * (ROW_TILE / 4), SF_LIN_THREADS);
    const size_t items = (size_t)bpp * c.n_pairs, cap = (size_t)a.num_sms * 2 * SF_LIN_BPS;  // resident blocks per SM, two rounds
    const unsigned grid = (unsigned)(items < cap ? items : cap);
    static const int max_stages = [] {  // A-B measurements: SF_LIN_UNSTAGED, SF_LIN_STAGES=2|3
        if (std::getenv("SF_LIN_UNSTAGED")) return 0;
        const char* e = std::getenv("SF_LIN_STAGES");
        const int v = e ? std::atoi(e) : LIN_MAX_STAGES;
        return v < 2 ? 2 : (v > LIN_MAX_STAGES ? LIN_MAX_STAGES : v);
    }();
    if (max_stages >= 3 && lin_dyn_smem(g.cols, 3) * SF_LIN_BPS <= LIN_MAX_STAGED_SMEM_PER_SM)
        linearise_kernel<3><<<grid, SF_LIN_THREADS, lin_dyn_smem(g.cols, 3), c.stream>>>(a, p, g, first, bpp);
    else if (max_stages >= 2 && lin_dyn_smem(g.cols, 2) * SF_LIN_BPS <= LIN_MAX_STAGED_SMEM_PER_SM)
        linearise_kernel<2><<<grid, SF_LIN_THREADS, lin_dyn_smem(g.cols, 2), c.stream>>>(a, p, g, first, bpp);
    else
        linearise_kernel<0><<<grid, SF_LIN_THREADS, 0, c.stream>>>(a, p, g, first, bpp);
    return 1;
}

void linearise_kernel_attrs() {  // attribute setup for the current device, outside stream capture (called by prepare_kernels in every sf_create)
    cudaFuncSetAttribute(linearise_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(LIN_MAX_STAGED_SMEM_PER_SM / SF_LIN_BPS));
    cudaFuncSetAttribute(linearise_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, 88);  // room for the stages of all resident blocks
    cudaFuncSetAttribute(linearise_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(LIN_MAX_STAGED_SMEM_PER_SM / SF_LIN_BPS));
    cudaFuncSetAttribute(linearise_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, 88);
}

int launch_step_prep(const Arena& a, const DevParams& p, int level_i, int k, const LaunchCfg& c) {
    step_prep_kernel<<<cdiv(c.n_pairs, 64), 64, 0, c.stream>>>(a, p, level_i, k, c.n_pairs);
    return 1;
}
}  // namespace sf
