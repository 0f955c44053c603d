// sf_k_irls.cu - IRLS passes, fused small-level IRLS loop, pose update, finish and the per-pixel weight image
// Part of the sm_100a kernels of the StaticFusion joint odometry + segmentation solver (launch interface: sf_kernels.cuh).
// One launch of each kernel serves the whole batch of frame pairs; data-dependent exits (IRLS convergence FrontEnd.cpp:679,
// outer-loop exit :1130, k-means :227) are per-pair flags in PairCtl that later launches test, so the host enqueues a static
// schedule with no synchronisation.  Compiled with -fmad=false: float expressions keep the reference's operation order and
// rounding; fused multiply-adds appear only where written explicitly.  Reference citations are relative to the upstream tree.
#include <cstdlib>

#include "sf_common.cuh"

namespace sf {

// K2: IRLS.  Both passes stream the raw rows written by linearise_kernel (14 floats + 1 label byte per pixel).
//
// Numerics (mirrors the oracle's EXACT policy): with m = 1/max pre-weight (FrontEnd.cpp:505-509),
//   res   = m * (-b_raw + sum_k Var_k * a_raw_k)                         (:644-646 on the raw row, then normalised)
//   w     = clamp(b_segm)/sqrt(1 + (res/(kc*aver_res))^2)                (:624-633)
//   aw_k  = (w * (m * 2^s_k)) * a_raw_k                                  (:628, scaled by the column's power of two)
// Normal equations as order-independent FIXED-POINT sums: the product of two scaled entries is exact in double; one
// double-precision FMA rounds it to a multiple of 2^-20 (ties to even) and adds it to an accumulator that stays inside the
// binade of 1.5 * 2^32, whose ulp is 2^-20 -- so every addition is exact and the accumulator's bit pattern, minus the
// constant's, is the integer sum of the rounded products.  Integer addition is associative, so the sums are
// bit-reproducible for any thread / block / GPU partition (sf_device.cuh).
__device__ __forceinline__ float residual_raw(const float* a, float b, const float* var) {
    float r = -b;
#pragma unroll
    for (int c = 0; c < 6; c++) r += var[c] * a[c];
    return r;
}

// exact warp sum of per-thread int64 partials (|sum| < 2^63): four 16-bit limbs, one REDUX each
__device__ __forceinline__ long long warp_sum_i64_exact(long long v) {
    const unsigned lo = (unsigned)v;
    const int hi = (int)(v >> 32);
    const unsigned l0 = __reduce_add_sync(0xffffffffu, lo & 0xffffu);
    const unsigned l1 = __reduce_add_sync(0xffffffffu, lo >> 16);
    const unsigned h0 = __reduce_add_sync(0xffffffffu, (unsigned)hi & 0xffffu);
    const int h1 = __reduce_add_sync(0xffffffffu, hi >> 16);
    return (long long)h1 * 281474976710656ll + ((long long)h0 << 32) + ((long long)l1 << 16) + (long long)l0;
}
// the integer a fixed-point accumulator holds
__device__ __forceinline__ long long fixacc_value(double acc) { return __double_as_longlong(acc) - QMAGIC_D_BITS; }

// ---- TMA bulk-copy pipeline -----------------------------------------------------------------------------------
// Every warp owns a ring of PS_STAGES tile buffers in shared memory.  Lane 0 arms the stage's mbarrier with the tile
// size and issues one cp.async.bulk (global -> shared, 3648 B); the warp waits on the barrier's phase parity, consumes
// the tile with conflict-free 8-byte shared loads (lane = 2 pixels) and refills the stage.  No block-wide
// synchronisation in the streaming loop; two blocks of 8 warps per SM keep 48 tiles (175 KB) in flight.
#ifndef SF_PS_WARPS
#define SF_PS_WARPS 8
#endif
#ifndef SF_PS_STAGES
#define SF_PS_STAGES 3
#endif
constexpr int PS_WARPS = SF_PS_WARPS;
constexpr int PS_STAGES = SF_PS_STAGES;
constexpr int PS_THREADS = PS_WARPS * 32;
constexpr int PS_BLOCKS_PER_SM = 2;
constexpr size_t PS_RING_BYTES = (size_t)PS_WARPS * PS_STAGES * TILE_BYTES;

// per-warp tile stream over the tiles [t0, t1) of one pair.  Warp w takes the tiles with (t - t0) % PS_WARPS == w, in
// an order that keeps consecutive tiles on the SAME image columns: `pattern` tiles span a whole number of image rows,
// so stepping by PS_WARPS * pattern tiles moves straight down; the remaining phases follow one after the other.
// (Cluster labels are vertically coherent, which lets pass 2 keep per-lane label runs in registers.)
struct TileStream {
    unsigned char* ring;          // this warp's PS_STAGES buffers
    unsigned long long* bars;     // this warp's PS_STAGES mbarriers
    unsigned phase;               // parity bit per stage, persists across items
    const unsigned char* src;     // first byte of the pair's tile array
    int count, issued;            // tiles to stream, tiles issued so far
    int it_ph, it_t, t_first, t_end, step, pattern;  // issue iterator (lane 0)
    int nw = PS_WARPS;            // warps of the block that share the tile range

    __device__ __forceinline__ void begin(const unsigned char* pair_tiles, int t0, int t1, int pat, int warp, int lane) {
        src = pair_tiles;
        t_first = t0 + warp; t_end = t1; pattern = pat; step = nw * pat;
        count = (t1 - t_first + nw - 1) / nw;
        if (count < 0) count = 0;
        issued = 0; it_ph = 0; it_t = t_first;
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the stages were last read through the generic proxy
            for (; issued < count && issued < PS_STAGES; issued++) issue(issued);
        }
        issued = __shfl_sync(0xffffffffu, issued, 0);
    }
    __device__ __forceinline__ void issue(int i) {  // lane 0 only; tiles are issued in traversal order
        while (it_t >= t_end) { it_ph++; it_t = t_first + it_ph * nw; }  // next phase: same warp slot, next column class
        const int st = i % PS_STAGES;
        mbar_arm(&bars[st], TILE_BYTES);
        bulk_load(ring + (size_t)st * TILE_BYTES, src + (size_t)it_t * TILE_BYTES, TILE_BYTES, &bars[st]);
        it_t += step;
    }
    __device__ __forceinline__ const unsigned char* wait(int i) {
        const int st = i % PS_STAGES;
        mbar_wait(&bars[st], (phase >> st) & 1u);
        phase ^= 1u << st;
        return ring + (size_t)st * TILE_BYTES;
    }
    __device__ __forceinline__ void release(int i, int lane) {  // the whole warp is done reading tile i
        __syncwarp();
        if (lane == 0 && i + PS_STAGES < count) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads before the async refill
            issue(i + PS_STAGES);
        }
    }
};

struct PassRing {
    unsigned char* ring;
    unsigned long long* bars;
};
template <int W = PS_WARPS>
__device__ __forceinline__ PassRing pass_ring_setup(unsigned char* dyn_smem, int warp, int lane, int tid) {
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(dyn_smem + (size_t)W * PS_STAGES * TILE_BYTES);
    if (tid < W * PS_STAGES) mbar_init(&bars[tid], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    PassRing r;
    r.ring = dyn_smem + (size_t)warp * PS_STAGES * TILE_BYTES;
    r.bars = bars + warp * PS_STAGES;
    return r;
}
constexpr size_t PS_DYN_SMEM = PS_RING_BYTES + PS_WARPS * PS_STAGES * sizeof(unsigned long long);

// ---- per-tile bodies shared by the multi-block passes and the fused per-pair kernel -------------------------------
// A lane takes the tile positions `lane` and `lane + 32` (sf_device.cuh: 32 pixels out of 64 consecutive ones of
// an image row per step; labels are spatially coherent, so pass 2's label groups are few): conflict-free 4-byte shared loads.
// pass 1 on one tile: robust weights (:615-637) and the fixed-point normal equations (:640-641) of 2 pixels per lane
__device__ __forceinline__ void pass1_tile(const unsigned char* tile, int lane, int it, float inv_max_c, float inv_max_d,
                                           float inv_c_Cauchy, const float* s_b, const float* var, const float* mc, const float* md,
                                           double (&acc)[27]) {
    const float* tr = reinterpret_cast<const float*>(tile);
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int px = lane + 32 * j;
        // invalid pixels carry zero rows (linearise_kernel) and get a zero weight: no branch, they add exactly 0
        const int vl = tile[TILE_ROW_BYTES + px];
        const float bw = (vl < NC) ? s_b[vl] : 0.f;
#pragma unroll
        for (int r = 0; r < 2; r++) {  // colour row, depth row
            float a[7];
#pragma unroll
            for (int k = 0; k < 7; k++) a[k] = tr[((r ? RW_AD : RW_AC) + k) * ROW_TILE + px];
            // res = -B before the first solve (:589), else A*Var - B (:644-646)
            const float res = (r ? inv_max_d : inv_max_c) * ((it == 1) ? -a[6] : residual_raw(a, a[6], var));
            const float w = bw * sqrt_rn_normal(rcp_rn_normal(1.f + sq(res * inv_c_Cauchy)));  // :627, :633
            double d[7];
#pragma unroll
            for (int k = 0; k < 7; k++) d[k] = (double)((w * (r ? md[k] : mc[k])) * a[k]);
            int q = 0;
#pragma unroll
            for (int ii = 0; ii < 6; ii++)
#pragma unroll
                for (int jj = ii; jj < 6; jj++) { acc[q] = fma(d[ii], d[jj], acc[q]); q++; }
#pragma unroll
            for (int ii = 0; ii < 6; ii++) acc[21 + ii] = fma(d[ii], d[6], acc[21 + ii]);
        }
    }
}

// pass 2 on one tile: residuals of the new solution (:644-646), |res|^2 and the per-label sums (:650-667)
__device__ __forceinline__ void pass2_tile(const unsigned char* tile, int lane, float inv_max_c, float inv_max_d, const float (&var)[6],
                                           float rscale, float lscale, long long* fix_w, int* cnt_w, double& rs) {
    const float* tr = reinterpret_cast<const float*>(tile);
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int px = lane + 32 * j;
        const int vl = tile[TILE_ROW_BYTES + px];
        const bool on = vl < NC;  // invalid pixels carry zero rows: their residual is 0
        float ac[7], ad[7];
#pragma unroll
        for (int k = 0; k < 7; k++) { ac[k] = tr[(RW_AC + k) * ROW_TILE + px]; ad[k] = tr[(RW_AD + k) * ROW_TILE + px]; }
        const float res_c = inv_max_c * residual_raw(ac, ac[6], var);
        const float res_d = inv_max_d * residual_raw(ad, ad[6], var);
        const float ress_here = fabsf(res_c) + fabsf(res_d);  // :660
        const double rc_s = (double)(res_c * rscale), rd_s = (double)(res_d * rscale);
        rs = fma(rc_s, rc_s, rs);
        rs = fma(rd_s, rd_s, rs);
        const unsigned q = (unsigned)__float2int_rn(ress_here * lscale);  // round(ress * 2^(rexp+19)) < 2^30
        // per-label sums: lanes are grouped by label (1-3 groups per warp), one REDUX per group
        unsigned todo = __ballot_sync(0xffffffffu, on);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            const int l = __shfl_sync(0xffffffffu, vl, leader);
            const bool mine = on && (vl == l);
            const unsigned grp = __ballot_sync(0xffffffffu, mine);
            const unsigned lo = __reduce_add_sync(0xffffffffu, mine ? (q & 0xffffu) : 0u);  // 32 terms < 2^30: two exact limbs
            const unsigned hi = __reduce_add_sync(0xffffffffu, mine ? (q >> 16) : 0u);
            if (lane == leader) { fix_w[l] += ((long long)hi << 16) + (long long)lo; cnt_w[l] += __popc(grp); }
            todo &= ~grp;
        }
        __syncwarp();
    }
}

// ---- per-pair tails, written against any state type with PairCtl's field names (global PairCtl or a shared copy) ----
// pass-1 tail (one thread): integer normal equations -> doubles, unpivoted LDL^T solve (:642), residual scale
template <class S>
__device__ __forceinline__ void irls_solve6(S& c, const long long* ne) {
    double AtA[36], F[36], AtB[6], x[6];
    unsigned char zero[6];
    int sx[7];
    for (int i = 0; i < 7; i++) sx[i] = c.sexp[i];
    int kk = 0;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = i; j < 6; j++) {
            const double vv = scale_pow2((double)ne[kk], -(sx[i] + sx[j]) - QFRAC_BITS);
            AtA[i * 6 + j] = vv; AtA[j * 6 + i] = vv; kk++;
        }
#pragma unroll
    for (int i = 0; i < 6; i++) AtB[i] = scale_pow2((double)ne[21 + i], -(sx[i] + sx[6]) - QFRAC_BITS);
#pragma unroll
    for (int i = 0; i < 36; i++) { F[i] = AtA[i]; c.AtA[i] = AtA[i]; }
    const int nz = ldlt_factor<6>(F, zero);
    ldlt_solve_factored<6>(F, zero, AtB, x);
    float rb = c.colbound[6];  // |res| <= |B| + sum_k |Var_k| |A_k|: scale of the integer |res|^2 sum
#pragma unroll
    for (int i = 0; i < 6; i++) {
        const float vi = (float)x[i];
        c.var[i] = vi;
        rb += fabsf(vi) * c.colbound[i];
    }
    c.rexp = scale_exponent(rb);
    if (nz) c.status |= SF_STATUS_SINGULAR;
}

// pass-2 tail (one warp): mean residuals (:666-667), 24x24 segmentation solve (SegmentationBackground.cpp:133-174),
// convergence test (:676-683).  lf / lc: this lane's label sum and count (lanes < 24); rs_total: integer |res|^2.
// Returns (in every lane) whether the IRLS loop of the pair has ended.
template <class S>
__device__ __forceinline__ bool irls_seg_tail(S& c, const DevParams& prm, long long lf, int lc, long long rs_total, const float (&var)[6], int it,
                                              double* s_A, double* s_rhs, double* s_x, unsigned char* s_zero, float* s_aver_label, int lane,
                                              float* ti) {
    const int N = c.n_valid;
    const long long tot = warp_sum_ll(lf);
    const float aver_res_old = c.aver_res;
    const int lsh = c.rexp + (LABEL_BITS - QSCALE_BITS - 1);  // scale of the per-label sums: |res_c| + |res_d| < 2^(11 - rexp)
    const float aver_new = (float)fixval(tot, lsh) / float(2 * N);  // :666
    if (lane < NC) s_aver_label[lane] = (float)fixval(lf, lsh) / float(2 * (lc + 1));  // :651,667 (counts start at 1)
    __syncwarp();
    if (prm.enable_segmentation) {
        // AtA_seg = diag(a^2) + (2 lambda_reg)^2 * Laplacian ; AtB_seg = a*B  
        const double aver = (double)aver_res_old;
        const double repr_res = (double)fmaxf(0.001f, aver_res_old);
        const double r0 = (double)prm.kb * repr_res / ((double)prm.kc_cauchy * aver);
        const double fixed_term = log(1.0 + r0 * r0);
        const double mult_res = 1.0 / ((double)prm.kc_cauchy * aver);
        const double wreg = 2.0 * (double)prm.lambda_reg;
        const double wreg2 = wreg * wreg;
        if (lane < NC) {
            const int l = lane;
            const unsigned row = c.conn[l] & ~(1u << l);
            for (int m = 0; m < NC; m++) {
                double lap = 0.0;
                if (m == l) lap = (double)__popc(row & 0xffffffu);
                else if (row & (1u << m)) lap = -1.0;
                s_A[l * 25 + m] = wreg2 * lap;
            }
            double aa, bb;
            const double ltw = (double)c.lambda_t_w[l];
            if (c.lambda_t_w[l] > 0.1f) {
                const double rl = (double)s_aver_label[l] * mult_res;
                const double dataterm = fixed_term - log(1.0 + rl * rl);
                aa = 2.0 * ltw * (double)prm.lambda_prior;
                bb = dataterm + 2.0 * (double)prm.lambda_prior * ltw * (double)c.b_prior[l];
            } else {
                aa = 2.0 * ltw;
                bb = 2.0 * ltw * (double)c.b_prior[l];
            }
            s_A[l * 25 + l] += aa * aa;
            s_rhs[l] = aa * bb;
        }
        __syncwarp();
        ldlt24_warp<25>(s_A, s_rhs, s_x, s_zero, lane);
        if (lane < NC) c.b_segm[lane] = (float)fmax(-1.0, fmin(2.0, s_x[lane]));
    }
    int done_i = 0;
    if (lane == 0) {
        const double rsq = scale_pow2((double)rs_total, -2 * c.rexp - QFRAC_BITS);
        c.res_sq = rsq;
        float delta = 0.f;  // :676
        for (int i = 0; i < 6; i++) { delta = fmaxf(delta, fabsf(c.prev_sol[i] - var[i])); c.prev_sol[i] = var[i]; }
        c.aver_res_old = aver_res_old;
        c.aver_res = aver_new;
        c.it_done = it;
        c.total_irls += 1;
        const bool done = (delta < prm.irls_delta_threshold) || (it == prm.max_iter_irls) || !(aver_new > 0.f);
        c.irls_done = done ? 1 : 0;
        done_i = done ? 1 : 0;
        if (ti) {
            for (int i = 0; i < 6; i++) ti[i] = var[i];
            ti[30] = aver_new; ti[31] = delta; ti[32] = (float)rsq;
        }
    }
    __syncwarp();
    if (lane < NC && ti) ti[6 + lane] = c.b_segm[lane];
    return __shfl_sync(0xffffffffu, done_i, 0) != 0;
}

__device__ __forceinline__ float* irls_trace_rec(const Arena& a, const DevParams& prm, int pair, int level_i, int k_outer, int it) {
    if (!a.trace || it > SF_TRACE_MAX_IRLS) return nullptr;
    return a.trace + ((size_t)pair * a.trace_steps + (level_i * prm.max_iter_per_level + k_outer)) * SF_TRACE_STEP + SF_TRACE_HDR +
           (it - 1) * SF_TRACE_IRLS;
}

// Work items of a pass launch: (pair still iterating, tile range).  The pairs come from the iteration's work list (odd it:
// iter_list0, even it: iter_list1; lengths in gcount[2], gcount[3]) and the number of items per pair is chosen on the device
// from the list length: about four items per resident block over the whole list, never fewer than 4 tiles per warp, a single
// item per pair when the list alone fills the GPU.  Any partition gives the same bits (integer sums).
struct PassItems {
    const int* list;
    int tiles_per_item, items_per_pair, total_items;
};
__device__ __forceinline__ PassItems pass_items(const Arena& a, const LevelGeom& g, int it, int resident_blocks) {
    PassItems p;
    const int par = (it - 1) & 1;
    p.list = par ? a.iter_list1 : a.iter_list0;
    const int n = a.gcount[2 + par];
    const int tiles = (int)tiles_per_pair((size_t)g.P);
    int want = n > 0 ? (4 * resident_blocks + n - 1) / n : 1;
    const int most = tiles / (4 * PS_WARPS) > 1 ? tiles / (4 * PS_WARPS) : 1;
    if (want > most) want = most;
    if (want < 1) want = 1;
    const int need = (tiles + MAX_TILES_PER_WARP_ITEM * PS_WARPS - 1) / (MAX_TILES_PER_WARP_ITEM * PS_WARPS);
    if (want < need) want = need;  // a thread's fixed-point accumulators stay inside their binade (sf_device.cuh)
    p.tiles_per_item = (tiles + want - 1) / want;
    p.items_per_pair = (tiles + p.tiles_per_item - 1) / p.tiles_per_item;
    p.total_items = p.items_per_pair * n;
    return p;
}

// pass 1: robust weights (:615-637), normal equations (:640-641), 6x6 solve (:642).
// Persistent blocks loop over the launch's work items.
__global__ void __launch_bounds__(PS_THREADS, PS_BLOCKS_PER_SM)
irls_pass1_kernel(Arena a, DevParams prm, LevelGeom g, int it, int resident_blocks, int pattern, int ctr_slot) {
    if (a.gcount[1] == 0) return;  // no pair is iterating any more (written by earlier kernels)
    const PassItems pi = pass_items(a, g, it, resident_blocks);
    const int tiles_per_item = pi.tiles_per_item, items_per_pair = pi.items_per_pair, total_items = pi.total_items;
    if (blockIdx.x == 0 && threadIdx.x == 0) a.gcount[2 + (it & 1)] = 0;  // pass 2 of this iteration appends the pairs that go on
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ float s_b[NC];
    __shared__ float s_var[6];
    __shared__ float s_mc[7], s_md[7];
    __shared__ long long s_part[PS_WARPS][28];
    __shared__ int s_last;
    const PassRing pr = pass_ring_setup(dyn_smem, warp, lane, tid);
    TileStream ts;
    ts.ring = pr.ring; ts.bars = pr.bars; ts.phase = 0;
    const int level_tiles = (int)tiles_per_pair((size_t)g.P);
    __shared__ int s_item;
    int item = blockIdx.x;  // first item is static, the rest is fetched from the launch's work counter (load balance)
    for (;; ) {
        if (item >= total_items) break;
        const int cur = item;
        __syncthreads();  // everyone has read the previous s_item / finished the previous item
        if (tid == 0) s_item = (int)gridDim.x + atomicAdd(&a.work_ctr[ctr_slot], 1);
        __syncthreads();
        item = s_item;
        const int slot = cur / items_per_pair, chunk = cur - slot * items_per_pair;
        const int pair = pi.list[slot];
        PairCtl& c = a.ctl[pair];
        if (!c.active || c.irls_done) continue;  // block-uniform
        const int t0 = chunk * tiles_per_item, t1 = min(t0 + tiles_per_item, level_tiles);
        ts.begin(a.tiles + (size_t)pair * tiles_per_pair(a.P0) * TILE_BYTES, t0, t1, pattern, warp, lane);  // copies fly while the constants load
        __syncthreads();  // shared state of the previous item is no longer read
        if (tid < NC) s_b[tid] = fmaxf(0.f, fminf(1.f, c.b_segm[tid]));  // :624
        if (tid < 6) s_var[tid] = c.var[tid];
        if (tid < 7) { s_mc[tid] = c.mcs[tid]; s_md[tid] = c.mds[tid]; }
        const float inv_max_c = c.inv_max_c, inv_max_d = c.inv_max_d;
        const float inv_c_Cauchy = 1.f / (prm.kc_cauchy * c.aver_res);  // :615
        __syncthreads();

        double acc[27];
#pragma unroll
        for (int i = 0; i < 27; i++) acc[i] = QMAGIC_D;
        for (int i = 0; i < ts.count; i++) {  // ts.count <= MAX_TILES_PER_WARP_ITEM (pass_items)
            const unsigned char* tile = ts.wait(i);
            pass1_tile(tile, lane, it, inv_max_c, inv_max_d, inv_c_Cauchy, s_b, s_var, s_mc, s_md, acc);  // per-pair constants stay in shared memory
            ts.release(i, lane);
        }
        // the accumulators' integers, reduced exactly and published with integer atomics
        long long mine = 0;
#pragma unroll
        for (int i = 0; i < 27; i++) {
            const long long ws = warp_sum_i64_exact(fixacc_value(acc[i]));
            if (lane == i) mine = ws;
        }
        if (lane < 27) s_part[warp][lane] = mine;
        __syncthreads();
        if (tid < 27) {
            long long t = 0;
#pragma unroll
            for (int w = 0; w < PS_WARPS; w++) t += s_part[w][tid];
            if (t) atomic_add_ll(&c.acc_ne[tid], t);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned t = atomicAdd(&c.ticket1, 1u);
            s_last = (t == (unsigned)items_per_pair - 1u) ? 1 : 0;
        }
        __syncthreads();
        if (!s_last) continue;
        __threadfence();
        if (tid == 0) {  // tail: one thread per pair solves the 6x6 system in double
            long long ne[27];
            for (int i = 0; i < 27; i++) ne[i] = __ldcg(&c.acc_ne[i]);
            irls_solve6(c, ne);
            for (int i = 0; i < 27; i++) c.acc_ne[i] = 0;
            c.ticket1 = 0;
        }
    }
}

// pass 2: residuals of the new solution (:644-646), per-label sums (:650-667), 24x24 segmentation
// solve (solveSegmIteration, SegmentationBackground.cpp:133-174), convergence test (:676-683)
__global__ void __launch_bounds__(PS_THREADS, PS_BLOCKS_PER_SM)
irls_pass2_kernel(Arena a, DevParams prm, LevelGeom g, int level_i, int k_outer, int it, int resident_blocks, int pattern, int ctr_slot) {
    if (a.gcount[1] == 0) return;
    const PassItems pi = pass_items(a, g, it, resident_blocks);
    const int tiles_per_item = pi.tiles_per_item, items_per_pair = pi.items_per_pair, total_items = pi.total_items;
    int* next_list = (it & 1) ? a.iter_list1 : a.iter_list0;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ long long s_fix[PS_WARPS][NC];  // per-warp label sums of round((|res_c|+|res_d|) * 2^(rexp+19)) < 2^30 each
    __shared__ int s_cnt[PS_WARPS][NC];
    __shared__ long long s_rs[PS_WARPS];
    __shared__ float s_var[6];
    __shared__ int s_last;
    __shared__ double s_A[NC * 25];
    __shared__ double s_rhs[NC], s_x[NC];
    __shared__ unsigned char s_zero[NC];
    __shared__ float s_aver_label[NC];
    const PassRing pr = pass_ring_setup(dyn_smem, warp, lane, tid);
    TileStream ts;
    ts.ring = pr.ring; ts.bars = pr.bars; ts.phase = 0;
    const int level_tiles = (int)tiles_per_pair((size_t)g.P);
    __shared__ int s_item;
    int item = blockIdx.x;  // first item is static, the rest is fetched from the launch's work counter (load balance)
    for (;; ) {
        if (item >= total_items) break;
        const int cur = item;
        __syncthreads();  // everyone has read the previous s_item / finished the previous item
        if (tid == 0) s_item = (int)gridDim.x + atomicAdd(&a.work_ctr[ctr_slot], 1);
        __syncthreads();
        item = s_item;
        const int slot = cur / items_per_pair, chunk = cur - slot * items_per_pair;
        const int pair = pi.list[slot];
        PairCtl& c = a.ctl[pair];
        if (!c.active || c.irls_done) continue;  // block-uniform
        const int t0 = chunk * tiles_per_item, t1 = min(t0 + tiles_per_item, level_tiles);
        ts.begin(a.tiles + (size_t)pair * tiles_per_pair(a.P0) * TILE_BYTES, t0, t1, pattern, warp, lane);
        __syncthreads();
        for (int q = tid; q < PS_WARPS * NC; q += PS_THREADS) { (&s_fix[0][0])[q] = 0; (&s_cnt[0][0])[q] = 0; }
        if (tid < 6) s_var[tid] = c.var[tid];
        const float inv_max_c = c.inv_max_c, inv_max_d = c.inv_max_d;
        const int rexp = c.rexp;
        const float rscale = ldexpf(1.f, rexp), lscale = ldexpf(1.f, rexp + (LABEL_BITS - QSCALE_BITS - 1));
        __syncthreads();
        float var[6];
#pragma unroll
        for (int i = 0; i < 6; i++) var[i] = s_var[i];

        double rs = QMAGIC_D;
        for (int i = 0; i < ts.count; i++) {
            const unsigned char* tile = ts.wait(i);
            pass2_tile(tile, lane, inv_max_c, inv_max_d, var, rscale, lscale, s_fix[warp], s_cnt[warp], rs);
            ts.release(i, lane);
        }
        const long long wrs = warp_sum_i64_exact(fixacc_value(rs));
        if (lane == 0) s_rs[warp] = wrs;
        __syncthreads();
        if (tid < NC) {
            long long f = 0;
            int n = 0;
#pragma unroll
            for (int w = 0; w < PS_WARPS; w++) { f += s_fix[w][tid]; n += s_cnt[w][tid]; }
            if (n) { atomic_add_ll(&c.lab_fix[tid], f); atomicAdd(&c.lab_cnt[tid], n); }
        }
        if (tid == 0) {
            long long t = 0;
            for (int w = 0; w < PS_WARPS; w++) t += s_rs[w];
            if (t) atomic_add_ll(&c.acc_rs, t);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned t = atomicAdd(&c.ticket2, 1u);
            s_last = (t == (unsigned)items_per_pair - 1u) ? 1 : 0;
        }
        __syncthreads();
        if (!s_last) continue;
        __threadfence();
        if (warp == 0) {
            // ---- tail, one warp ----
            long long lf = 0;
            int lc = 0;
            if (lane < NC) { lf = __ldcg(&c.lab_fix[lane]); lc = __ldcg(&c.lab_cnt[lane]); }
            const long long rs_total = __ldcg(&c.acc_rs);
            const bool done = irls_seg_tail(c, prm, lf, lc, rs_total, var, it, s_A, s_rhs, s_x, s_zero, s_aver_label, lane,
                                            irls_trace_rec(a, prm, pair, level_i, k_outer, it));
            if (lane == 0) {
                c.acc_rs = 0;
                c.ticket2 = 0;
                if (done) atomicSub(&a.gcount[1], 1);
                else next_list[atomicAdd(&a.gcount[2 + (it & 1)], 1)] = pair;  // next iteration's work list
            }
            if (lane < NC) { c.lab_fix[lane] = 0; c.lab_cnt[lane] = 0; }
        }
    }
}

// ---- queue-driven IRLS loop: ONE launch runs the IRLS loops of all pairs of a large level ----------------------------
// The static schedule used to enqueue max_iter_irls x (pass 1, pass 2) launches per step; after the second iteration most of
// them found no pair (the drivers' threshold ends the loop after 2-3 iterations for ~95 % of the pairs) and every launch
// boundary made the pairs that were still iterating wait for each other.  Here persistent blocks pop WORK ITEMS
// (pair, pass, tile range) from a queue in global memory; the block that finishes the last item of a pair's pass runs that
// pass's tail (6x6 solve / 24x24 segmentation solve + exit test) and pushes the items of the pair's next pass, or retires the
// pair.  Pairs advance through their iterations independently; there is no grid-wide barrier and no wait other than for a
// queue slot to be filled, and a slot a block waits for is always filled by a block that is running (a pair cannot retire
// while one of its items is unprocessed), so the loop cannot deadlock whatever the number of resident blocks.
// A slot is valid when it carries the step's generation (Arena::gcount[4], bumped by step_begin_kernel): no reset pass.
// Per-pair state written by one block and read by another goes through L2: every block executes a device-scope fence after
// it pops an item (which also drops the SM's L1 lines) before it touches the pair.
constexpr int LOOP_BLOCKS_PER_PAIR = 2;  // measured 2 / 4 / 8 / 16 / 64: 5.35 / 5.38 / 5.37 / 5.40 / 5.43 ms per step (3 lanes), flat with one lane
__device__ __forceinline__ unsigned long long q_load(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int vol_load(const int* p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// items a pass of one pair is cut into, from the number of pairs still iterating (same rule as pass_items)
__device__ __forceinline__ int loop_items_per_pair(int n_iterating, int level_tiles, int resident_blocks) {
    int want = n_iterating > 0 ? (4 * resident_blocks + n_iterating - 1) / n_iterating : 1;
    const int most = level_tiles / (4 * PS_WARPS) > 1 ? level_tiles / (4 * PS_WARPS) : 1;
    if (want > most) want = most;
    if (want < 1) want = 1;
    const int need = (level_tiles + MAX_TILES_PER_WARP_ITEM * PS_WARPS - 1) / (MAX_TILES_PER_WARP_ITEM * PS_WARPS);
    if (want < need) want = need;
    const int per = (level_tiles + want - 1) / want;
    return (level_tiles + per - 1) / per;
}
// one thread: publish the items of `pass` (0 = pass 1, 1 = pass 2) of a pair.  The pair's state was written before.
__device__ __forceinline__ void loop_push(const Arena& a, PairCtl& c, int pair, int pass, int level_tiles, int resident_blocks, unsigned gen) {
    const int K = loop_items_per_pair(vol_load(&a.gcount[1]), level_tiles, resident_blocks);
    c.items_cur = K;
    __threadfence();  // state (and items_cur) before the items
    const int base = atomicAdd(&a.gcount[6], K);
    for (int j = 0; j < K; j++)
        if (base + j < a.q_cap)
            atomicExch(&a.irls_q[base + j], ((unsigned long long)gen << 32) | ((unsigned long long)pair << 12) | ((unsigned long long)pass << 11) | (unsigned long long)j);
}

__global__ void __launch_bounds__(PS_THREADS, PS_BLOCKS_PER_SM)
irls_loop_kernel(Arena a, DevParams prm, LevelGeom g, int level_i, int k_outer, int resident_blocks, int pattern, int blocks_per_pair) {
    if (a.gcount[1] == 0) return;  // no pair enters the loop in this step (written by step_prep_kernel, an earlier launch)
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ float s_b[NC];
    __shared__ float s_var[6];
    __shared__ float s_mc[7], s_md[7];
    __shared__ long long s_part[PS_WARPS][28];
    __shared__ long long s_fix[PS_WARPS][NC];
    __shared__ int s_cnt[PS_WARPS][NC];
    __shared__ long long s_rs[PS_WARPS];
    __shared__ double s_A[NC * 25];
    __shared__ double s_rhs[NC], s_x[NC];
    __shared__ unsigned char s_zero[NC];
    __shared__ float s_aver_label[NC];
    __shared__ int s_last;
    __shared__ long long s_item;
    const PassRing pr = pass_ring_setup(dyn_smem, warp, lane, tid);
    TileStream ts;
    ts.ring = pr.ring; ts.bars = pr.bars; ts.phase = 0;
    const int level_tiles = (int)tiles_per_pair((size_t)g.P);
    const unsigned gen = (unsigned)a.gcount[4];
    if (tid == 0) atomicAdd(&a.gcount[7], 1);  // blocks alive in this loop
    if (blockIdx.x == 0) {  // the pairs step_prep_kernel listed enter the loop: items of their first pass 1
        const int n = a.gcount[2];
        for (int s = tid; s < n; s += PS_THREADS) {
            const int pair = a.iter_list0[s];
            loop_push(a, a.ctl[pair], pair, 0, level_tiles, resident_blocks, gen);
        }
    }
    for (;;) {
        // ---- pop: claim the head slot once it carries this step's generation.  A block that finds nothing to do leaves when
        // every pair has retired, or when more blocks are alive than the pairs still iterating can keep busy (LOOP_BLOCKS_PER_PAIR
        // each): an idle block would only keep its SM from the kernels of the other lanes.
        __syncthreads();  // the previous item's shared state is no longer read
        if (tid == 0) {
            long long got = -1;
            for (long long spin = 0; spin < (1ll << 24) && got == -1; spin++) {
                if (vol_load(&a.gcount[6]) > vol_load(&a.gcount[5])) {  // items are waiting: take a ticket, then its slot
                    const int idx = atomicAdd(&a.gcount[5], 1);
                    if (idx >= a.q_cap) break;
                    for (long long w = 0; w < (1ll << 24); w++) {  // (another block may have taken the last one: then this slot is filled by the next push)
                        const unsigned long long v = q_load(&a.irls_q[idx]);
                        if ((unsigned)(v >> 32) == gen) { got = (long long)(v & 0xffffffffull); break; }
                        if (vol_load(&a.gcount[1]) == 0) { got = -3; break; }
                        __nanosleep(64);
                    }
                    break;  // got == -1: the wait for the ticket's slot ran out
                }
                const int rem = vol_load(&a.gcount[1]);
                if (rem == 0) { got = -3; break; }  // every pair has retired: nothing will be pushed any more
                const int live = vol_load(&a.gcount[7]);
                if (live > blocks_per_pair * rem) {
                    if (atomicCAS(&a.gcount[7], live, live - 1) == live) got = -2;
                    continue;
                }
                __nanosleep(200);
            }
            if (got == -1) {  // the bounded wait ran out: must never happen; flag the batch and release the other blocks
                const int n = a.gcount[2];
                for (int q = 0; q < n; q++) atomicOr(&a.ctl[a.iter_list0[q]].status, SF_STATUS_INTERNAL);
                atomicExch(&a.gcount[1], 0);
            }
            if (got != -2 && got < 0) atomicSub(&a.gcount[7], 1);
            s_item = got;
        }
        __syncthreads();
        const long long item = s_item;
        if (item < 0) break;
        __threadfence();  // acquire: the pusher's state is visible, stale L1 lines are dropped
        const int pair = (int)(item >> 12), pass = (int)((item >> 11) & 1), chunk = (int)(item & 2047);
        PairCtl& c = a.ctl[pair];
        const int K = c.items_cur;
        const int tiles_per_item = (level_tiles + K - 1) / K;
        const int t0 = chunk * tiles_per_item, t1 = min(t0 + tiles_per_item, level_tiles);
        const int it = c.it_done + 1;
        ts.begin(a.tiles + (size_t)pair * tiles_per_pair(a.P0) * TILE_BYTES, t0, t1, pattern, warp, lane);  // copies fly while the constants load
        const float inv_max_c = c.inv_max_c, inv_max_d = c.inv_max_d;
        if (pass == 0) {
            // ---------------- pass 1: robust weights (:615-637), normal equations (:640-641)
            if (tid < NC) s_b[tid] = fmaxf(0.f, fminf(1.f, c.b_segm[tid]));  // :624
            if (tid < 6) s_var[tid] = c.var[tid];
            if (tid < 7) { s_mc[tid] = c.mcs[tid]; s_md[tid] = c.mds[tid]; }
            const float inv_c_Cauchy = 1.f / (prm.kc_cauchy * c.aver_res);  // :615
            __syncthreads();
            double acc[27];
#pragma unroll
            for (int i = 0; i < 27; i++) acc[i] = QMAGIC_D;
            for (int i = 0; i < ts.count; i++) {
                const unsigned char* tile = ts.wait(i);
                pass1_tile(tile, lane, it, inv_max_c, inv_max_d, inv_c_Cauchy, s_b, s_var, s_mc, s_md, acc);
                ts.release(i, lane);
            }
            long long mine = 0;
#pragma unroll
            for (int i = 0; i < 27; i++) {
                const long long ws = warp_sum_i64_exact(fixacc_value(acc[i]));
                if (lane == i) mine = ws;
            }
            if (lane < 27) s_part[warp][lane] = mine;
            __syncthreads();
            if (tid < 27) {
                long long t = 0;
#pragma unroll
                for (int w = 0; w < PS_WARPS; w++) t += s_part[w][tid];
                if (t) atomic_add_ll(&c.acc_ne[tid], t);
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const unsigned t = atomicAdd(&c.ticket1, 1u);
                s_last = (t == (unsigned)K - 1u) ? 1 : 0;
            }
            __syncthreads();
            if (!s_last) continue;
            __threadfence();
            if (tid == 0) {  // tail: 6x6 solve in double (:642), then the pair's pass 2 enters the queue
                long long ne[27];
                for (int i = 0; i < 27; i++) ne[i] = __ldcg(&c.acc_ne[i]);
                irls_solve6(c, ne);
                for (int i = 0; i < 27; i++) c.acc_ne[i] = 0;
                c.ticket1 = 0;
                loop_push(a, c, pair, 1, level_tiles, resident_blocks, gen);
            }
        } else {
            // ---------------- pass 2: residuals of the new solution (:644-646), per-label sums (:650-667)
            for (int q = tid; q < PS_WARPS * NC; q += PS_THREADS) { (&s_fix[0][0])[q] = 0; (&s_cnt[0][0])[q] = 0; }
            if (tid < 6) s_var[tid] = c.var[tid];
            const int rexp = c.rexp;
            const float rscale = ldexpf(1.f, rexp), lscale = ldexpf(1.f, rexp + (LABEL_BITS - QSCALE_BITS - 1));
            __syncthreads();
            float var[6];
#pragma unroll
            for (int i = 0; i < 6; i++) var[i] = s_var[i];
            double rs = QMAGIC_D;
            for (int i = 0; i < ts.count; i++) {
                const unsigned char* tile = ts.wait(i);
                pass2_tile(tile, lane, inv_max_c, inv_max_d, var, rscale, lscale, s_fix[warp], s_cnt[warp], rs);
                ts.release(i, lane);
            }
            const long long wrs = warp_sum_i64_exact(fixacc_value(rs));
            if (lane == 0) s_rs[warp] = wrs;
            __syncthreads();
            if (tid < NC) {
                long long f = 0;
                int n = 0;
#pragma unroll
                for (int w = 0; w < PS_WARPS; w++) { f += s_fix[w][tid]; n += s_cnt[w][tid]; }
                if (n) { atomic_add_ll(&c.lab_fix[tid], f); atomicAdd(&c.lab_cnt[tid], n); }
            }
            if (tid == 0) {
                long long t = 0;
                for (int w = 0; w < PS_WARPS; w++) t += s_rs[w];
                if (t) atomic_add_ll(&c.acc_rs, t);
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const unsigned t = atomicAdd(&c.ticket2, 1u);
                s_last = (t == (unsigned)K - 1u) ? 1 : 0;
            }
            __syncthreads();
            if (!s_last) continue;
            __threadfence();
            if (warp == 0) {  // tail, one warp: segmentation solve (SegmentationBackground.cpp:133-174), exit test (:676-683)
                long long lf = 0;
                int lc = 0;
                if (lane < NC) { lf = __ldcg(&c.lab_fix[lane]); lc = __ldcg(&c.lab_cnt[lane]); }
                const long long rs_total = __ldcg(&c.acc_rs);
                const bool done = irls_seg_tail(c, prm, lf, lc, rs_total, var, it, s_A, s_rhs, s_x, s_zero, s_aver_label, lane,
                                                irls_trace_rec(a, prm, pair, level_i, k_outer, it));
                if (lane < NC) { c.lab_fix[lane] = 0; c.lab_cnt[lane] = 0; }
                __threadfence();  // every lane's part of the pair's state (b_segm, cleared cells) is out before lane 0 publishes
                __syncwarp();
                if (lane == 0) {
                    c.acc_rs = 0;
                    c.ticket2 = 0;
                    if (done) { __threadfence(); atomicSub(&a.gcount[1], 1); }  // the pair retires: its state is complete before the count drops
                    else loop_push(a, c, pair, 0, level_tiles, resident_blocks, gen);
                }
            }
        }
    }
}

// ---- fused IRLS loop: ONE block runs all iterations of a pair (both passes, both solves, the exit test) with the
// per-pair state in shared memory.  Used for the levels whose per-pair data is small: there the multi-block passes
// are bound by launch / ticket / tail latency (12 launches per step), not by bandwidth.  Same per-tile bodies and
// tails as the passes above, so the result is bit-identical (all sums are integers).
struct FusedState {  // the PairCtl fields the tails touch
    float var[6], prev_sol[6];
    double AtA[36];
    double res_sq;
    float b_segm[NC], b_prior[NC], lambda_t_w[NC];
    unsigned conn[NC];
    float colbound[7];
    int sexp[7];
    int rexp, n_valid;
    float aver_res, aver_res_old;
    int it_done, total_irls, irls_done, status;
};

template <int W, int BPS>
__global__ void __launch_bounds__(W * 32, BPS)
irls_fused_kernel(Arena a, DevParams prm, LevelGeom g, int level_i, int k_outer, int n_pairs, int pattern, int ctr_slot) {
    if (a.gcount[1] == 0) return;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ FusedState st;
    __shared__ float s_b[NC];
    __shared__ float s_mc[7], s_md[7];
    __shared__ long long s_part[W][28];
    __shared__ long long s_ne[27];
    __shared__ long long s_fix[W][NC];
    __shared__ int s_cnt[W][NC];
    __shared__ long long s_rs[W];
    __shared__ double s_A[NC * 25];
    __shared__ double s_rhs[NC], s_x[NC];
    __shared__ unsigned char s_zero[NC];
    __shared__ float s_aver_label[NC];
    __shared__ int s_next, s_done;
    const PassRing pr = pass_ring_setup<W>(dyn_smem, warp, lane, tid);
    TileStream ts;
    ts.ring = pr.ring; ts.bars = pr.bars; ts.phase = 0; ts.nw = W;
    const int level_tiles = (int)tiles_per_pair((size_t)g.P);
    int pair = blockIdx.x;
    for (;; ) {
        if (pair >= n_pairs) break;
        const int cur = pair;
        __syncthreads();
        if (tid == 0) s_next = (int)gridDim.x + atomicAdd(&a.work_ctr[ctr_slot], 1);
        __syncthreads();
        pair = s_next;
        PairCtl& c = a.ctl[cur];
        if (!c.active || c.irls_done) continue;  // block-uniform (degenerate steps have irls_done == 2)
        const unsigned char* pair_tiles = a.tiles + (size_t)cur * tiles_per_pair(a.P0) * TILE_BYTES;
        ts.begin(pair_tiles, 0, level_tiles, pattern, warp, lane);  // first tiles fly while the state loads
        if (tid < NC) { st.b_segm[tid] = c.b_segm[tid]; st.b_prior[tid] = c.b_prior[tid]; st.lambda_t_w[tid] = c.lambda_t_w[tid]; st.conn[tid] = c.conn[tid]; }
        if (tid < 6) { st.var[tid] = c.var[tid]; st.prev_sol[tid] = c.prev_sol[tid]; }
        if (tid < 7) { st.colbound[tid] = c.colbound[tid]; st.sexp[tid] = c.sexp[tid]; s_mc[tid] = c.mcs[tid]; s_md[tid] = c.mds[tid]; }
        if (tid == 0) {
            st.rexp = c.rexp; st.n_valid = c.n_valid; st.aver_res = c.aver_res; st.aver_res_old = c.aver_res_old;
            st.it_done = c.it_done; st.total_irls = c.total_irls; st.irls_done = 0; st.status = c.status; st.res_sq = 0.0;
        }
        const float inv_max_c = c.inv_max_c, inv_max_d = c.inv_max_d;
        __syncthreads();
        for (int it = 1; it <= prm.max_iter_irls; it++) {
            // ---------------- pass 1 ----------------
            if (it > 1) ts.begin(pair_tiles, 0, level_tiles, pattern, warp, lane);
            if (tid < NC) s_b[tid] = fmaxf(0.f, fminf(1.f, st.b_segm[tid]));  // :624
            const float inv_c_Cauchy = 1.f / (prm.kc_cauchy * st.aver_res);  // :615
            __syncthreads();
            double acc[27];
#pragma unroll
            for (int i = 0; i < 27; i++) acc[i] = QMAGIC_D;
            for (int i = 0; i < ts.count; i++) {  // level_tiles / W <= MAX_TILES_PER_WARP_ITEM (fused_level_ok)
                const unsigned char* tile = ts.wait(i);
                pass1_tile(tile, lane, it, inv_max_c, inv_max_d, inv_c_Cauchy, s_b, st.var, s_mc, s_md, acc);
                ts.release(i, lane);
            }
            ts.begin(pair_tiles, 0, level_tiles, pattern, warp, lane);  // pass 2's first tiles fly during the reduction and the solve
            long long mine = 0;
#pragma unroll
            for (int i = 0; i < 27; i++) {
                const long long ws = warp_sum_i64_exact(fixacc_value(acc[i]));
                if (lane == i) mine = ws;
            }
            if (lane < 27) s_part[warp][lane] = mine;
            __syncthreads();
            if (tid < 27) {
                long long t = 0;
#pragma unroll
                for (int w = 0; w < W; w++) t += s_part[w][tid];
                s_ne[tid] = t;
            }
            for (int q = tid; q < W * NC; q += (W * 32)) { (&s_fix[0][0])[q] = 0; (&s_cnt[0][0])[q] = 0; }
            __syncthreads();
            if (tid == 0) irls_solve6(st, s_ne);
            __syncthreads();
            // ---------------- pass 2 ----------------
            float var[6];
#pragma unroll
            for (int i = 0; i < 6; i++) var[i] = st.var[i];
            const int rexp = st.rexp;
            const float rscale = ldexpf(1.f, rexp), lscale = ldexpf(1.f, rexp + (LABEL_BITS - QSCALE_BITS - 1));
            double rs = QMAGIC_D;
            for (int i = 0; i < ts.count; i++) {
                const unsigned char* tile = ts.wait(i);
                pass2_tile(tile, lane, inv_max_c, inv_max_d, var, rscale, lscale, s_fix[warp], s_cnt[warp], rs);
                ts.release(i, lane);
            }
            const long long wrs = warp_sum_i64_exact(fixacc_value(rs));
            if (lane == 0) s_rs[warp] = wrs;
            __syncthreads();
            if (warp == 0) {
                long long lf = 0;
                int lc = 0;
                if (lane < NC)
#pragma unroll
                    for (int w = 0; w < W; w++) { lf += s_fix[w][lane]; lc += s_cnt[w][lane]; }
                long long rs_total = 0;
                for (int w = 0; w < W; w++) rs_total += s_rs[w];
                const bool done = irls_seg_tail(st, prm, lf, lc, rs_total, var, it, s_A, s_rhs, s_x, s_zero, s_aver_label, lane,
                                                irls_trace_rec(a, prm, cur, level_i, k_outer, it));
                if (lane == 0) s_done = done ? 1 : 0;
            }
            __syncthreads();
            if (s_done) break;
        }
        // ---------------- write the state back ----------------
        if (tid < NC) c.b_segm[tid] = st.b_segm[tid];
        if (tid < 6) { c.var[tid] = st.var[tid]; c.prev_sol[tid] = st.prev_sol[tid]; }
        if (tid >= 32 && tid < 32 + 36) c.AtA[tid - 32] = st.AtA[tid - 32];
        if (tid == 0) {
            c.rexp = st.rexp; c.res_sq = st.res_sq; c.aver_res = st.aver_res; c.aver_res_old = st.aver_res_old;
            c.it_done = st.it_done; c.total_irls = st.total_irls; c.irls_done = 1; c.status = st.status;
            atomicSub(&a.gcount[1], 1);
        }
    }
}

// ------------------------------------------------------------------------------------------
// D1: covariance, motion filter, SE(3) update, outer-loop exit (FrontEnd.cpp:689, 713-772, 1130)
// ------------------------------------------------------------------------------------------
__global__ void pose_update_kernel(Arena a, DevParams prm, int level_i, int k, int n_pairs) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= n_pairs) return;
    PairCtl& c = a.ctl[pair];
    if (!c.active) return;
    float* tr = a.trace ? a.trace + ((size_t)pair * a.trace_steps + (level_i * prm.max_iter_per_level + k)) * SF_TRACE_STEP : nullptr;
    if (c.irls_done == 2) {  // degenerate step: no estimate, pose untouched
        for (int i = 0; i < 6; i++) c.twist_level[i] = 0.f;
    } else {
        double tw[6], Tod[16];
        for (int i = 0; i < 6; i++) tw[i] = (double)c.var[i];
        for (int i = 0; i < 16; i++) Tod[i] = (double)c.T[i];
        if (prm.use_motion_filter) {
            double F[36], cov[36], ev[6], V[36];
            unsigned char zero[6];
            for (int i = 0; i < 36; i++) F[i] = c.AtA[i];
            ldlt_factor<6>(F, zero);
            const double res_sq = c.res_sq;
            for (int cc = 0; cc < 6; cc++) {  // est_cov = AtA^-1 * ||res||^2, :689
                double e[6] = {0, 0, 0, 0, 0, 0}, x[6];
                e[cc] = 1.0;
                ldlt_solve_factored<6>(F, zero, e, x);
                for (int r = 0; r < 6; r++) cov[r * 6 + cc] = x[r] * res_sq;
            }
            for (int i = 0; i < 6; i++)
                for (int j = 0; j < i; j++) { const double m = 0.5 * (cov[i * 6 + j] + cov[j * 6 + i]); cov[i * 6 + j] = m; cov[j * 6 + i] = m; }
            jacobi_eig6(cov, ev, V);
            double kai_b[6], kai_b_old[6], kai_loc_sub[6], lg[6];
            se3_log(Tod, lg);  // :736-738
            for (int i = 0; i < 6; i++) kai_loc_sub[i] = (double)c.twist_old[i] - lg[i];
            for (int i = 0; i < 6; i++) {
                double s1 = 0, s2 = 0;
                for (int q = 0; q < 6; q++) { s1 += V[q * 6 + i] * tw[q]; s2 += V[q * 6 + i] * kai_loc_sub[q]; }
                kai_b[i] = s1; kai_b_old[i] = s2;
            }
            const float e = prm.exp_neg_level[level_i];  // expf(-level), :745
            const double cf = (double)(prm.previous_speed_eig_weight * e), df = (double)(prm.previous_speed_const_weight * e);
            double fil[6];
            for (int i = 0; i < 6; i++) fil[i] = (kai_b[i] + (cf * ev[i] + df) * kai_b_old[i]) / (1.0 + cf * ev[i] + df);  // :750
            for (int i = 0; i < 6; i++) {
                double s = 0;
                for (int q = 0; q < 6; q++) s += V[i * 6 + q] * fil[q];
                tw[i] = s;
            }
        }
        for (int i = 0; i < 6; i++) { c.twist_level[i] = (float)tw[i]; tw[i] = (double)c.twist_level[i]; }
        double E[16], Tn[16];
        se3_exp(tw, E);  // :759-766
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) {
                double s = 0;
                for (int q = 0; q < 4; q++) s += E[i * 4 + q] * Tod[q * 4 + j];
                Tn[i * 4 + j] = s;
            }
        for (int i = 0; i < 16; i++) { c.T[i] = (float)Tn[i]; Tn[i] = (double)c.T[i]; }
        double lg[6];
        se3_log(Tn, lg);  // :769-771
        for (int i = 0; i < 6; i++) c.twist_odom[i] = (float)lg[i];
        rigid_inverse(c.T, c.Tinv);
    }
    double nrm = 0;
    for (int i = 0; i < 6; i++) nrm += (double)c.twist_level[i] * (double)c.twist_level[i];
    if (sqrt(nrm) < (double)prm.outer_exit_threshold) c.break_level = level_i;  // :1130
    {
        int* st = a.stepstat + ((size_t)pair * a.trace_steps + (level_i * prm.max_iter_per_level + k)) * 2;
        st[0] = c.n_valid; st[1] = c.it_done;
    }
    if (tr) {
        tr[4] = (float)c.it_done;
        for (int i = 0; i < 6; i++) { tr[56 + i] = c.twist_level[i]; tr[79 + i] = c.twist_odom[i]; }
        for (int i = 0; i < 16; i++) tr[62 + i] = c.T[i];
        tr[78] = (float)c.status;
    }
}

// end of runSolver (FrontEnd.cpp:1139-1144) + result block
__global__ void finish_kernel(Arena a, int n_pairs) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= n_pairs) return;
    PairCtl& c = a.ctl[pair];
    double R[9], Ri[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[i * 3 + j] = (double)c.T[i * 4 + j];
    const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
    const double id = 1.0 / det;
    Ri[0] = (R[4] * R[8] - R[5] * R[7]) * id; Ri[1] = (R[2] * R[7] - R[1] * R[8]) * id; Ri[2] = (R[1] * R[5] - R[2] * R[4]) * id;
    Ri[3] = (R[5] * R[6] - R[3] * R[8]) * id; Ri[4] = (R[0] * R[8] - R[2] * R[6]) * id; Ri[5] = (R[2] * R[3] - R[0] * R[5]) * id;
    Ri[6] = (R[3] * R[7] - R[4] * R[6]) * id; Ri[7] = (R[1] * R[6] - R[0] * R[7]) * id; Ri[8] = (R[0] * R[4] - R[1] * R[3]) * id;
    PairOut& o = a.out[pair];
    for (int h = 0; h < 2; h++)
        for (int i = 0; i < 3; i++) {
            float s = 0.f;
            for (int j = 0; j < 3; j++) s += (float)Ri[i * 3 + j] * c.twist_odom[3 * h + j];
            o.twist_old[3 * h + i] = s;
        }
    for (int i = 0; i < 16; i++) o.T[i] = c.T[i];
    for (int l = 0; l < NC; l++) o.b_segm[l] = c.b_segm[l];
    o.irls_iters = c.total_irls;
    o.status = c.status;
}

// K7: per-pixel static weight (buildSegmImage, SegmentationBackground.cpp:176-197); row-major output.
// perClusterAverageResidual is NaN unless the 5-frame history ran (FrontEnd.cpp:105); NaN < 0.017 is false.
// 4 pixels per thread (one uchar4 label load, one float4 store), a block covers up to 8192 pixels of a pair so that the
// 24-entry table is built once per 8 K pixels instead of once per 256
constexpr int SEGM_PIXELS_PER_BLOCK = 8192;
__global__ void __launch_bounds__(256) segm_image_kernel(Arena a, LevelGeom g0) {
    const int pair = blockIdx.y;
    __shared__ float s_b[NC + 1];
    if (threadIdx.x < NC) {
        float b = fmaxf(0.f, fminf(1.f, a.ctl[pair].b_segm[threadIdx.x]));
        if ((double)a.pcar[pair * NC + threadIdx.x] < 0.017) b = fmaxf(b, 1.0f - b);  // :190-194 (double literal)
        s_b[threadIdx.x] = b;
    }
    if (threadIdx.x == NC) s_b[NC] = 1.f;  // :181-185 invalid cluster = static
    __syncthreads();
    const uchar4* lab4 = reinterpret_cast<const uchar4*>(a.labels + (size_t)pair * a.pyr_stride + g0.off);  // cols % 4 == 0 on every level
    float4* out4 = reinterpret_cast<float4*>(a.b_perpixel + (size_t)pair * a.P0);
    const int c0 = blockIdx.x * (SEGM_PIXELS_PER_BLOCK / 4), c1 = min(c0 + SEGM_PIXELS_PER_BLOCK / 4, g0.P >> 2);
    for (int ch = c0 + threadIdx.x; ch < c1; ch += 256) {
        const uchar4 l = __ldg(lab4 + ch);
        out4[ch] = make_float4(s_b[l.x], s_b[l.y], s_b[l.z], s_b[l.w]);
    }
}

int irls_chunk_iters(int P) {
    int it = (P + 1024 * 24 - 1) / (1024 * 24);
    if (it < 1) it = 1;
    if (it > 8) it = 8;
    return it;
}

// tiles that span a whole number of image rows: lcm(cols, ROW_TILE) / ROW_TILE
static inline int tile_pattern(int cols) {
    int a = cols, b = ROW_TILE;
    while (b) { const int t = a % b; a = b; b = t; }
    return cols / a;
}
// Fused-kernel shapes: many pairs -> small blocks (4 warps, 5 per SM) so that every pair of the batch is resident at once and
// the serial solves of one pair hide behind the streaming of the others; few pairs -> 12 warps per pair.
constexpr int FW_SMALL = 4, FB_SMALL = 4;
constexpr size_t fused_dyn_smem(int w) { return (size_t)w * PS_STAGES * TILE_BYTES + (size_t)w * PS_STAGES * sizeof(unsigned long long); }
void linearise_kernel_attrs();  // sf_k_linearise.cu
void prepare_kernels() { extern void pass_kernel_attrs_impl(); pass_kernel_attrs_impl(); linearise_kernel_attrs(); }
void pass_kernel_attrs_impl() {  // per device (function attributes are not shared between devices): called by every sf_create
    cudaFuncSetAttribute(irls_pass1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PS_DYN_SMEM);
    cudaFuncSetAttribute(irls_pass2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PS_DYN_SMEM);
    cudaFuncSetAttribute(irls_loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PS_DYN_SMEM);
    cudaFuncSetAttribute(irls_fused_kernel<PS_WARPS, PS_BLOCKS_PER_SM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_dyn_smem(PS_WARPS));
    cudaFuncSetAttribute(irls_fused_kernel<FW_SMALL, FB_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_dyn_smem(FW_SMALL));
}

// grid of a pass launch: the resident blocks, or fewer when even the finest partition of every pair has fewer items
// (measured: a quarter-size grid for the late, mostly empty iterations costs the same - an empty launch is launch-bound)
static inline int pass_grid(const Arena& a, int P, int n_pairs) {
    const int cap = a.num_sms * PS_BLOCKS_PER_SM;
    const int tiles = (int)tiles_per_pair((size_t)P);
    const int most = tiles / (4 * PS_WARPS) > 1 ? tiles / (4 * PS_WARPS) : 1;
    const long long total = (long long)most * n_pairs;
    return total < cap ? (int)total : cap;
}
int launch_irls_pass1(const Arena& a, const DevParams& p, const LevelGeom& g, int, int, int it, const LaunchCfg& c) {
    const int slot = (*c.next_ctr)++ % MAX_WORK_CTRS;
    irls_pass1_kernel<<<pass_grid(a, g.P, c.n_pairs), PS_THREADS, PS_DYN_SMEM, c.stream>>>(a, p, g, it, a.num_sms * PS_BLOCKS_PER_SM, tile_pattern(g.cols), slot);
    return 1;
}

// items one IRLS loop of a lane can push: every pass of every iteration of every pair, cut as finely as loop_items_per_pair allows
size_t irls_queue_capacity(int max_pairs, int max_iter_irls, size_t P0) {
    const size_t tiles = tiles_per_pair(P0);
    const size_t most = tiles / (4 * PS_WARPS) > 1 ? tiles / (4 * PS_WARPS) : 1;
    const size_t need = (tiles + (size_t)MAX_TILES_PER_WARP_ITEM * PS_WARPS - 1) / ((size_t)MAX_TILES_PER_WARP_ITEM * PS_WARPS);
    return (size_t)max_pairs * (size_t)max_iter_irls * 2 * ((most > need ? most : need) + 1) + 64;
}

// whole IRLS loop of a step for a large level: persistent blocks, device-side item queue
int launch_irls_loop(const Arena& a, const DevParams& p, const LevelGeom& g, int level_i, int k, const LaunchCfg& c) {
    const int resident = a.num_sms * PS_BLOCKS_PER_SM;
    static const int bpp = std::getenv("SF_LOOP_BPP") ? std::atoi(std::getenv("SF_LOOP_BPP")) : LOOP_BLOCKS_PER_PAIR;  // A-B measurements
    irls_loop_kernel<<<pass_grid(a, g.P, c.n_pairs), PS_THREADS, PS_DYN_SMEM, c.stream>>>(a, p, g, level_i, k, resident, tile_pattern(g.cols), bpp);
    return 1;
}

// whole IRLS loop of a step, one block per pair (for the levels where irls_fused_level() says so)
int launch_irls_fused(const Arena& a, const DevParams& p, const LevelGeom& g, int level_i, int k, const LaunchCfg& c) {
    const int slot = (*c.next_ctr)++ % MAX_WORK_CTRS;
    if (c.n_pairs > 2 * a.num_sms) {
        const int cap = a.num_sms * FB_SMALL;
        irls_fused_kernel<FW_SMALL, FB_SMALL><<<c.n_pairs < cap ? c.n_pairs : cap, FW_SMALL * 32, fused_dyn_smem(FW_SMALL), c.stream>>>(
            a, p, g, level_i, k, c.n_pairs, tile_pattern(g.cols), slot);
    } else {
        const int cap = a.num_sms * PS_BLOCKS_PER_SM;
        irls_fused_kernel<PS_WARPS, PS_BLOCKS_PER_SM><<<c.n_pairs < cap ? c.n_pairs : cap, PS_THREADS, fused_dyn_smem(PS_WARPS), c.stream>>>(
            a, p, g, level_i, k, c.n_pairs, tile_pattern(g.cols), slot);
    }
    return 1;
}

int launch_irls_pass2(const Arena& a, const DevParams& p, const LevelGeom& g, int level_i, int k, int it, const LaunchCfg& c) {
    const int slot = (*c.next_ctr)++ % MAX_WORK_CTRS;
    irls_pass2_kernel<<<pass_grid(a, g.P, c.n_pairs), PS_THREADS, PS_DYN_SMEM, c.stream>>>(a, p, g, level_i, k, it, a.num_sms * PS_BLOCKS_PER_SM, tile_pattern(g.cols), slot);
    return 1;
}

int launch_pose_update(const Arena& a, const DevParams& p, int level_i, int k, const LaunchCfg& c) {
    pose_update_kernel<<<cdiv(c.n_pairs, 32), 32, 0, c.stream>>>(a, p, level_i, k, c.n_pairs);
    return 1;
}

int launch_finish(const Arena& a, const DevParams&, const LevelGeom& g0, const LaunchCfg& c) {
    finish_kernel<<<cdiv(c.n_pairs, 64), 64, 0, c.stream>>>(a, c.n_pairs);
    return 1;
}

int launch_segm_image(const Arena& a, const LevelGeom& g0, const LaunchCfg& c) {
    segm_image_kernel<<<dim3(cdiv(g0.P, SEGM_PIXELS_PER_BLOCK), c.n_pairs), 256, 0, c.stream>>>(a, g0);
    return 1;
}
}  // namespace sf
