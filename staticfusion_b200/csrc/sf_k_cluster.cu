// sf_k_cluster.cu - geometric clustering: k-means at half resolution, full-resolution labelling + adjacency, label pyramid (KMeans.cpp)
// Part of the sm_100a kernels of the StaticFusion joint odometry + segmentation solver (launch interface: sf_kernels.cuh).
// One launch of each kernel serves the whole batch of frame pairs; data-dependent exits (IRLS convergence FrontEnd.cpp:679,
// outer-loop exit :1130, k-means :227) are per-pair flags in PairCtl that later launches test, so the host enqueues a static
// schedule with no synchronisation.  Compiled with -fmad=false: float expressions keep the reference's operation order and
// rounding; fused multiply-adds appear only where written explicitly.  Reference citations are relative to the upstream tree.
#include <cstdlib>

#include "sf_common.cuh"

namespace sf {

// ------------------------------------------------------------------------------------------
// K6: geometric clustering (KMeans.cpp)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned float_order_key(float x) {
    const unsigned b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_from_order_key(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Per-centre candidate lists (KMeans.cpp:172-183 / :249-259): row l lists all centres sorted by their squared distance
// to centre l; an entry carries the candidate's coordinates and that distance in ONE 16-byte word, so the pruned search
// below costs one shared load per candidate.  Stable rank sort == std::stable_sort by distance (the reference's
// std::sort differs only for exactly equal distances).
struct CandTable {
    // (z, x, y of the candidate, squared centre-to-centre distance).  Rows are padded to 25 entries: with 24 (384 bytes = 96 words)
    // every row starts in the same bank and lanes looking at different clusters serialise; 400 bytes shift a row by 4 banks
    float4 cand[NC][NC + 1];
    unsigned char tidx[NC][NC];   // the candidate's cluster index
};

// nearest-centre search with the reference's pruned traversal (KMeans.cpp:192-212 / :267-289)
__device__ __forceinline__ int nearest_pruned(int last_label, float p0, float p1, float p2, const float4* cen4, const CandTable& t) {
    const float4 cl = cen4[last_label];
    const float distance_to_last_label = sqnorm3(cl.x, cl.y, cl.z, p0, p1, p2);
    float best_distance = distance_to_last_label;
    int best_li = 0;
    const float lim = 4.f * distance_to_last_label;
    for (int li = 1; li < NC; ++li) {
        const float4 cc = t.cand[last_label][li];
        if (cc.w > lim) break;
        const float distance_to_label = sqnorm3(cc.x, cc.y, cc.z, p0, p1, p2);
        if (distance_to_label < best_distance) { best_distance = distance_to_label; best_li = li; }
    }
    return best_li ? (int)t.tidx[last_label][best_li] : last_label;
}

// build the table with the whole block: distances first, then a stable rank sort
// (rank = number of entries that are smaller, or equal with a smaller index)
__device__ __forceinline__ void build_cand_table_block(const float4* cen4, CandTable& t, float (*scratch)[NC], int tid, int nthreads) {
    for (int i = tid; i < NC * NC; i += nthreads) {
        const int l = i / NC, li = i - l * NC;
        scratch[l][li] = sqnorm3(cen4[l].x, cen4[l].y, cen4[l].z, cen4[li].x, cen4[li].y, cen4[li].z);
    }
    __syncthreads();
    for (int i = tid; i < NC * NC; i += nthreads) {
        const int l = i / NC, li = i - l * NC;
        const float dv = scratch[l][li];
        int rank = 0;
        for (int j = 0; j < NC; j++) {
            const float dj = scratch[l][j];
            rank += (dj < dv || (dj == dv && j < li)) ? 1 : 0;
        }
        t.cand[l][rank] = make_float4(cen4[li].x, cen4[li].y, cen4[li].z, dv);
        t.tidx[l][rank] = (unsigned char)li;
    }
    __syncthreads();
}

// rebuild the table from the sorted lists published by kmeans_kernel
__device__ __forceinline__ void load_cand_table_block(const PairCtl& c, float4* cen4, CandTable& t, int tid, int nthreads) {
    if (tid < NC) cen4[tid] = make_float4(c.kmeans[tid], c.kmeans[NC + tid], c.kmeans[2 * NC + tid], 0.f);
    __syncthreads();
    for (int i = tid; i < NC * NC; i += nthreads) {
        const int li = c.tbl_idx[i];
        t.cand[i / NC][i % NC] = make_float4(cen4[li].x, cen4[li].y, cen4[li].z, c.tbl_dist[i]);
        (&t.tidx[0][0])[i] = (unsigned char)li;
    }
    __syncthreads();
}

// one block per pair: seeds + medians (initializeKMeans, KMeans.cpp:63-135) and the Lloyd iterations at level 1
// (kMeans3DCoord, KMeans.cpp:167-228).  Every thread owns a contiguous range of 4-pixel chunks, so labels form long
// runs that are accumulated in registers and flushed on a label change; centre sums are fixed-point integers, so
// the result does not depend on the traversal order.  The seed labelling (nearest seed in pixel space, KMeans.cpp:87-101)
// depends on the image size only: Arena::seed_map holds it, computed once per context.
constexpr int KM_THREADS = 512;
constexpr int KM_WARPS = KM_THREADS / 32;
#ifndef SF_KM_QUEUE
#define SF_KM_QUEUE 0  // 1 = warp-queue compaction of the candidate loop (bit-identical, measured no faster: the kernel is latency- and barrier-bound)
#endif
#ifndef SF_KM_BPS
#define SF_KM_BPS 4
#endif
#if SF_KM_QUEUE
constexpr int KM_QUEUE = 64;        // per-warp relabel queue: up to 31 waiting + 32 new entries
#endif
constexpr int KM_LIST_CAP = 3072;  // relabelled pixels one Lloyd iteration can record before it falls back to a full re-accumulation
struct KmWarpBins {
    long long w0[KM_WARPS][NC], w1[KM_WARPS][NC], w2[KM_WARPS][NC];
    int wn[KM_WARPS][NC];
};
struct KmLloydSmem {  // live during the Lloyd iterations; shares its storage with the median histograms
    CandTable t;
    union {
        KmWarpBins b;                 // full accumulation: per-warp private bins
        unsigned list[KM_LIST_CAP];   // incremental update: (pixel << 10) | (old label << 5) | new label of every relabelled pixel
    };
};
union KmSmem {
    int hist[NC][256];
    KmLloydSmem l;
};
__global__ void __launch_bounds__(KM_THREADS, SF_KM_BPS) kmeans_kernel(Arena a, DevParams prm, LevelGeom g1) {
    const int pair = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const int frame = a.cur_idx[pair];
    const float* depth = a.pyr_d + (size_t)frame * a.pyr_stride + g1.off;
    uint8_t* labels = a.labels + (size_t)pair * a.pyr_stride + g1.off;
    PairCtl& c = a.ctl[pair];

    __shared__ __align__(16) KmSmem sm;
    __shared__ unsigned prefix[NC];
    __shared__ int rank[NC], csize[NC];
    __shared__ float4 cen[NC];
    __shared__ float scratch[NC][NC];
    __shared__ long long sums0[NC], sums1[NC], sums2[NC];
    __shared__ int cnt[NC];
    __shared__ int s_conv;
    const int warp = tid >> 5;

    if (tid < NC) csize[tid] = 0;
    __syncthreads();
    const int nchunks = g1.P >> 2;  // every level has cols % 4 == 0
    const int per = (nchunks + KM_THREADS - 1) / KM_THREADS;
    const int c0 = min(tid * per, nchunks), c1 = min(c0 + per, nchunks);
    const float4* depth4 = reinterpret_cast<const float4*>(depth);
    uchar4* labels4 = reinterpret_cast<uchar4*>(labels);
    const uchar4* seed4 = reinterpret_cast<const uchar4*>(a.seed_map);

    // seed labels (KMeans.cpp:87-101)
    {
        int run_lab = -1, run_n = 0;
        for (int ch = c0; ch < c1; ch++) {
            const float4 z4 = depth4[ch];
            const uchar4 s4 = __ldg(seed4 + ch);
            const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
            const int ss[4] = {s4.x, s4.y, s4.z, s4.w};
            unsigned char out[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int lab = (zz[j] != 0.f) ? ss[j] : (int)LABEL_NONE;
                if (lab != LABEL_NONE) {
                    if (lab != run_lab) {
                        if (run_n) atomicAdd(&csize[run_lab], run_n);
                        run_lab = lab; run_n = 0;
                    }
                    run_n++;
                }
                out[j] = (unsigned char)lab;
            }
            labels4[ch] = make_uchar4(out[0], out[1], out[2], out[3]);
        }
        if (run_n) atomicAdd(&csize[run_lab], run_n);
    }
    __syncthreads();
    // per-cluster median = element of rank size/2 (nth_element, KMeans.cpp:118-125): 4-pass radix select
    if (tid < NC) { prefix[tid] = 0; rank[tid] = csize[tid] / 2; }
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 24 - 8 * pass;
        for (int i = tid; i < NC * 256; i += KM_THREADS) (&sm.hist[0][0])[i] = 0;
        __syncthreads();
        int run_bin = -1, run_n = 0;
        for (int ch = c0; ch < c1; ch++) {
            const float4 z4 = depth4[ch];
            const uchar4 l4 = labels4[ch];
            const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
            const int ll[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int l = ll[j];
                if (l != LABEL_NONE) {
                    const unsigned key = float_order_key(zz[j]);
                    if (pass == 0 || (key >> (shift + 8)) == prefix[l]) {
                        const int bin = (l << 8) | (int)((key >> shift) & 255u);
                        if (bin != run_bin) {
                            if (run_n) atomicAdd(&(&sm.hist[0][0])[run_bin], run_n);
                            run_bin = bin; run_n = 0;
                        }
                        run_n++;
                    }
                }
            }
        }
        if (run_n) atomicAdd(&(&sm.hist[0][0])[run_bin], run_n);
        __syncthreads();
        for (int cl = warp; cl < NC; cl += KM_WARPS) {  // first bin whose running count exceeds the rank: 8 bins per lane, warp scan
            if (csize[cl] <= 0) continue;  // warp-uniform
            int h[8], mine = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) { h[q] = sm.hist[cl][lane * 8 + q]; mine += h[q]; }
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int r = rank[cl];  // < number of keys that carry the prefix = the warp's total
            const unsigned over = __ballot_sync(0xffffffffu, incl > r);
            const int owner = over ? __ffs(over) - 1 : 31;
            __syncwarp();
            if (lane == owner) {
                int rr = r - (incl - mine), b = 0;
                while (b < 7 && rr >= h[b]) { rr -= h[b]; b++; }
                rank[cl] = rr;
                prefix[cl] = (prefix[cl] << 8) | (unsigned)(lane * 8 + b);
            }
        }
        __syncthreads();
    }
    if (tid < NC) {  // KMeans.cpp:116-134
        if (csize[tid] > 0) {
            const float z = float_from_order_key(prefix[tid]);
            cen[tid] = make_float4(z, (prm.km_u_label[tid] - g1.disp_u) * z * g1.inv_f, (prm.km_v_label[tid] - g1.disp_v) * z * g1.inv_f, 0.f);
        } else {
            cen[tid] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __syncthreads();
    // Lloyd iterations (iter_kmeans - 1 = 9, KMeans.cpp:142,167).  The centre sums are integers, so they can be kept across
    // iterations and updated with the pixels that changed their label only (subtract from the old cluster, add to the new
    // one): the result is the same integer as the reference's from-scratch sum.  An iteration relabels every pixel and
    // records the changes in a shared list; the first iteration, and any iteration with more than KM_LIST_CAP changes,
    // re-accumulates all pixels instead (run-length accumulation in registers, per-warp private bins).
    KmLloydSmem& L = sm.l;
    __shared__ int s_list_n;
#if SF_KM_QUEUE
    __shared__ float4 s_queue[KM_WARPS][KM_QUEUE];  // per-warp queue of pixels waiting for the candidate loop: (z, x, y, pixel << 5 | label)
#endif
    __shared__ int s_dl[NC][10];  // limb sums of the incremental update: 3 coordinates x 3 limbs + count
    int its_done = 0;
    for (int it = 0; it < 9; it++) {
        its_done = it + 1;
        build_cand_table_block(cen, L.t, scratch, tid, KM_THREADS);
        if (tid == 0) s_list_n = 0;
        for (int i = tid; i < NC * 10; i += KM_THREADS) (&s_dl[0][0])[i] = 0;
        __syncthreads();
        const bool record = it > 0;
#if SF_KM_QUEUE
        // relabel (KMeans.cpp:187-217).  Most pixels leave the pruned search before its first candidate (the nearest other
        // centre is more than twice as far from their centre as they are); the others take 1-6 candidates.  To keep the
        // candidate loop from running at the slowest lane's trip count with most lanes idle, a pixel that needs the loop is
        // put in the warp's queue and the loop runs on 32 queued pixels at a time, one per lane.
        {
            float4* wq = s_queue[warp];
            int q_head = 0, q_count = 0;
            auto run_queue = [&](int n) {  // lanes < n take one queued pixel each
                if (lane < n) {
                    const float4 e = wq[(q_head + lane) & (KM_QUEUE - 1)];
                    const unsigned code = __float_as_uint(e.w);
                    const int old = (int)(code & 31u), pix = (int)(code >> 5);
                    const int lab = nearest_pruned(old, e.x, e.y, e.z, cen, L.t);
                    if (lab != old) {
                        labels[pix] = (uint8_t)lab;
                        if (record) {
                            const int slot = atomicAdd(&s_list_n, 1);
                            if (slot < KM_LIST_CAP) L.list[slot] = ((unsigned)pix << 10) | ((unsigned)old << 5) | (unsigned)lab;
                        }
                    }
                }
                __syncwarp();  // every lane has read its entry: the slots may be refilled
            };
            for (int k = 0; k < per; k++) {  // same trip count for every lane: the queue operations are warp-collective
                const int ch = c0 + k;
                const bool act = ch < c1;
                float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                uchar4 l4 = make_uchar4(0, 0, 0, 0);
                if (act) { z4 = depth4[ch]; l4 = labels4[ch]; }
                const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
                const int ll[4] = {l4.x, l4.y, l4.z, l4.w};
                const int p0 = ch << 2;
                int v, u0;
                split_rc(p0, g1, v, u0);
                const float cy = g1.inv_f * (float(v) - g1.disp_v);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float z = zz[j];
                    bool need = false;
                    float x = 0.f, y = 0.f;
                    if (z != 0.f) {  // inactive lanes carry z == 0
                        x = (g1.inv_f * (float(u0 + j) - g1.disp_u)) * z;  // xxPyr, FrontEnd.cpp:386
                        y = cy * z;
                        const float4 cl = cen[ll[j]];
                        const float d_own = sqnorm3(cl.x, cl.y, cl.z, z, x, y);
                        need = !(L.t.cand[ll[j]][1].w > 4.f * d_own);  // the loop of nearest_pruned would not break at once
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, need);
                    if (m) {  // warp-uniform
                        if (need) {
                            const int pos = (q_head + q_count + __popc(m & ((1u << lane) - 1u))) & (KM_QUEUE - 1);
                            wq[pos] = make_float4(z, x, y, __uint_as_float(((unsigned)(p0 + j) << 5) | (unsigned)ll[j]));
                        }
                        q_count += __popc(m);
                        __syncwarp();
                        if (q_count >= 32) { run_queue(32); q_head = (q_head + 32) & (KM_QUEUE - 1); q_count -= 32; }
                    }
                }
            }
            run_queue(q_count);
        }
#else
        // relabel (KMeans.cpp:187-217).  A warp takes 16 x 8 pixel tiles (lane = 4-pixel chunk lx of row ly): its lanes then look
        // at 1-3 clusters, so the centre / candidate loads are shared-memory broadcasts and the lanes' trip counts are similar
        // (a lane per far-apart pixel range costs ~4 wavefronts per load and runs every lane at the slowest lane's trip count)
        {
            const int cpr = g1.cols >> 2;  // 4-pixel chunks per row
            const int tiles_x = (cpr + 3) >> 2, tiles_y = (g1.rows + 7) >> 3;
            const int lx = lane & 3, ly = lane >> 2;
            for (int tile = warp; tile < tiles_x * tiles_y; tile += KM_WARPS) {
                const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
                const int v = ty * 8 + ly, cx = tx * 4 + lx;
                if (v >= g1.rows || cx >= cpr) continue;
                const int ch = v * cpr + cx;
                const float4 z4 = depth4[ch];
                const uchar4 l4 = labels4[ch];
                const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
                int ll[4] = {l4.x, l4.y, l4.z, l4.w};
                const int p0 = ch << 2, u0 = cx << 2;
                const float cy = g1.inv_f * (float(v) - g1.disp_v);
                bool any = false;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float z = zz[j];
                    if (z != 0.f) {
                        const float x = (g1.inv_f * (float(u0 + j) - g1.disp_u)) * z;  // xxPyr, FrontEnd.cpp:386
                        const float y = cy * z;
                        const int old = ll[j];
                        const int lab = nearest_pruned(old, z, x, y, cen, L.t);
                        if (lab != old) {
                            ll[j] = lab;
                            any = true;
                            if (record) {
                                const int slot = atomicAdd(&s_list_n, 1);
                                if (slot < KM_LIST_CAP) L.list[slot] = ((unsigned)(p0 + j) << 10) | ((unsigned)old << 5) | (unsigned)lab;
                            }
                        }
                    }
                }
                if (any) labels4[ch] = make_uchar4((unsigned char)ll[0], (unsigned char)ll[1], (unsigned char)ll[2], (unsigned char)ll[3]);
            }
        }
#endif
        __syncthreads();
        const int n_changed = s_list_n;
        bool incremental = false;  // block-uniform
        if (record && n_changed <= KM_LIST_CAP) {
            // incremental update of the sums from the list (KMeans.cpp:213-216 restricted to the pixels that moved).  A 64-bit
            // fixed-point term is cut into 16-bit limbs and every limb is added with a native 32-bit shared atomic (64-bit shared
            // atomics are CAS loops); at most KM_LIST_CAP * 65535 < 2^31 per cell; the limbs are recombined per cluster below.
            for (int i = tid; i < n_changed; i += KM_THREADS) {  // s_dl was cleared before the relabel pass
                const unsigned e = L.list[i];
                const int pix = (int)(e >> 10);
                const int lab_old = (int)((e >> 5) & 31u), lab_new = (int)(e & 31u);
                const float z = depth[pix];
                int v, u;
                split_rc(pix, g1, v, u);
                const float x = (g1.inv_f * (float(u) - g1.disp_u)) * z;
                const float y = (g1.inv_f * (float(v) - g1.disp_v)) * z;
                const long long q[3] = {fixq(z, FIX_KMEANS), fixq(x, FIX_KMEANS), fixq(y, FIX_KMEANS)};
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const int lo = (int)(q[k] & 0xffff), mid = (int)((q[k] >> 16) & 0xffff), hi = (int)(q[k] >> 32);  // q = hi 2^32 + mid 2^16 + lo
                    atomicAdd(&s_dl[lab_new][3 * k], lo); atomicAdd(&s_dl[lab_new][3 * k + 1], mid); atomicAdd(&s_dl[lab_new][3 * k + 2], hi);
                    atomicAdd(&s_dl[lab_old][3 * k], -lo); atomicAdd(&s_dl[lab_old][3 * k + 1], -mid); atomicAdd(&s_dl[lab_old][3 * k + 2], -hi);
                }
                atomicAdd(&s_dl[lab_new][9], 1); atomicAdd(&s_dl[lab_old][9], -1);
            }
            incremental = true;
        } else {
            // full accumulation of the relabelled level (KMeans.cpp:213-216)
            for (int i = tid; i < KM_WARPS * NC; i += KM_THREADS) { (&L.b.w0[0][0])[i] = 0; (&L.b.w1[0][0])[i] = 0; (&L.b.w2[0][0])[i] = 0; (&L.b.wn[0][0])[i] = 0; }
            __syncthreads();
            int run_lab = 0, run_n = 0;
            long long r0 = 0, r1 = 0, r2 = 0;
            for (int k = 0; k < per; k++) {  // same trip count for every lane: the flush below is warp-collective
                const int ch = c0 + k;
                const bool act = ch < c1;
                float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                uchar4 l4 = make_uchar4(0, 0, 0, 0);
                if (act) { z4 = depth4[ch]; l4 = labels4[ch]; }
                const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
                const int ll[4] = {l4.x, l4.y, l4.z, l4.w};
                const int p0 = ch << 2;
                int v, u0;
            split_rc(p0, g1, v, u0);
                const float cy = g1.inv_f * (float(v) - g1.disp_v);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float z = zz[j];
                    const bool on = act && (z != 0.f);
                    const int lab = on ? ll[j] : run_lab;
                    const bool change = on && (lab != run_lab) && (run_n > 0);
                    warp_serial_flush(change, run_lab, r0, r1, r2, run_n, 0, L.b.w0[warp], L.b.w1[warp], L.b.w2[warp], L.b.wn[warp], nullptr, lane);
                    if (on) {
                        const float x = (g1.inv_f * (float(u0 + j) - g1.disp_u)) * z;  // xxPyr, FrontEnd.cpp:386
                        const float y = cy * z;
                        if (lab != run_lab) { run_lab = lab; run_n = 0; r0 = 0; r1 = 0; r2 = 0; }
                        r0 += fixq(z, FIX_KMEANS); r1 += fixq(x, FIX_KMEANS); r2 += fixq(y, FIX_KMEANS);
                        run_n++;
                    }
                }
            }
            warp_serial_flush(run_n > 0, run_lab, r0, r1, r2, run_n, 0, L.b.w0[warp], L.b.w1[warp], L.b.w2[warp], L.b.wn[warp], nullptr, lane);
        }
        __syncthreads();
        if (warp == 0) {  // one warp closes the iteration: cluster totals, new centres, convergence test (no block barrier in between)
            float m = 0.f;
            if (lane < NC) {
                if (incremental) {
                    const int* d = s_dl[lane];
                    sums0[lane] += (long long)d[0] + ((long long)d[1] << 16) + ((long long)d[2] << 32);
                    sums1[lane] += (long long)d[3] + ((long long)d[4] << 16) + ((long long)d[5] << 32);
                    sums2[lane] += (long long)d[6] + ((long long)d[7] << 16) + ((long long)d[8] << 32);
                    cnt[lane] += d[9];
                } else {
                    long long a0 = 0, a1 = 0, a2 = 0;
                    int n = 0;
                    for (int w = 0; w < KM_WARPS; w++) { a0 += L.b.w0[w][lane]; a1 += L.b.w1[w][lane]; a2 += L.b.w2[w][lane]; n += L.b.wn[w][lane]; }
                    sums0[lane] = a0; sums1[lane] = a1; sums2[lane] = a2; cnt[lane] = n;
                }
                const int n = cnt[lane];  // KMeans.cpp:219-221 (empty clusters collapse to the origin)
                const float4 nb = make_float4(n > 0 ? (float)(fixval(sums0[lane], FIX_KMEANS) / (double)n) : 0.f,
                                              n > 0 ? (float)(fixval(sums1[lane], FIX_KMEANS) / (double)n) : 0.f,
                                              n > 0 ? (float)(fixval(sums2[lane], FIX_KMEANS) / (double)n) : 0.f, 0.f);
                // KMeans.cpp:224-227: max |old - new| over the 72 coordinates
                m = fmaxf(0.f, fmaxf(fmaxf(fabsf(cen[lane].x - nb.x), fabsf(cen[lane].y - nb.y)), fabsf(cen[lane].z - nb.z)));
                cen[lane] = nb;
            }
            const unsigned mb = __reduce_max_sync(0xffffffffu, __float_as_uint(m));  // non-negative floats order like their bit patterns
            if (lane == 0) s_conv = (__uint_as_float(mb) < 1e-2f) ? 1 : 0;
        }
        __syncthreads();
        if (s_conv) break;  // s_conv is next written after the barriers of the following iteration
    }
    // publish centres + the final sorted table for the full-resolution labelling (KMeans.cpp:232-259)
    build_cand_table_block(cen, L.t, scratch, tid, KM_THREADS);
    if (tid < NC) { c.kmeans[tid] = cen[tid].x; c.kmeans[NC + tid] = cen[tid].y; c.kmeans[2 * NC + tid] = cen[tid].z; }
    if (tid == 0) c.km_iters = its_done;
    for (int i = tid; i < NC * NC; i += KM_THREADS) { c.tbl_dist[i] = L.t.cand[i / NC][i % NC].w; c.tbl_idx[i] = (&L.t.tidx[0][0])[i]; }
}

// Full-resolution labelling (KMeans.cpp:263-291) fused with the cluster adjacency (computeRegionConnectivity,
// KMeans.cpp:297-341).  A block takes a band of image rows of one pair: it labels the band plus the first row of the
// next band (the adjacency test looks one row down; that row is labelled twice rather than exchanged), keeps depth and
// labels of the band in shared memory, then runs the adjacency test on them.  One table load serves the whole band.
constexpr int LB_THREADS = 256;
constexpr int LB_SMEM_PIXELS = 6144;  // depth + label of the band's pixels: 5 bytes each (30 KB)
__global__ void __launch_bounds__(LB_THREADS) label_connect_kernel(Arena a, DevParams prm, LevelGeom g0, LevelGeom g1, int band_rows, int bands_per_pair) {
    const int pair = blockIdx.x / bands_per_pair, band = blockIdx.x - pair * bands_per_pair;
    const int tid = threadIdx.x;
    const int frame = a.cur_idx[pair];
    PairCtl& c = a.ctl[pair];
    __shared__ float4 cen[NC];
    __shared__ __align__(16) CandTable t;
    __shared__ unsigned conn[NC];
    __shared__ __align__(16) float s_z[LB_SMEM_PIXELS];
    __shared__ __align__(16) unsigned char s_l[LB_SMEM_PIXELS];
    if (tid < NC) conn[tid] = 0;
    load_cand_table_block(c, cen, t, tid, LB_THREADS);
    const int r0 = band * band_rows, r1 = min(r0 + band_rows, g0.rows);
    const int rl = min(r1 + 1, g0.rows);  // rows labelled here (one look-ahead row)
    const float* depth = a.pyr_d + (size_t)frame * a.pyr_stride + g0.off;
    uint8_t* lab0 = a.labels + (size_t)pair * a.pyr_stride + g0.off;
    const uint8_t* lab1 = a.labels + (size_t)pair * a.pyr_stride + g1.off;
    const int cpr = g0.cols >> 2;  // 4-pixel chunks per row
    for (int ch = tid; ch < (rl - r0) * cpr; ch += LB_THREADS) {
        int rr, u0;
        split_rc(ch << 2, g0, rr, u0);  // cols % 4 == 0: chunk ch starts at pixel 4 ch of the band
        const int v = r0 + rr;
        const float4 z4 = ldg4(depth + (size_t)v * g0.cols + u0);
        const uchar2 low = *reinterpret_cast<const uchar2*>(lab1 + (size_t)(v >> 1) * g1.cols + (u0 >> 1));
        const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
        const int lw[4] = {low.x, low.x, low.y, low.y};
        unsigned char out[4];
        const float cy = g0.inv_f * (float(v) - g0.disp_v);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            out[j] = LABEL_NONE;
            if (zz[j] != 0.f) {
                const int last_label = (lw[j] == LABEL_NONE) ? 0 : lw[j];
                const float x = (g0.inv_f * (float(u0 + j) - g0.disp_u)) * zz[j];
                const float y = cy * zz[j];
                out[j] = (unsigned char)nearest_pruned(last_label, zz[j], x, y, cen, t);
            }
        }
        const uchar4 o4 = make_uchar4(out[0], out[1], out[2], out[3]);
        *reinterpret_cast<float4*>(s_z + rr * g0.cols + u0) = z4;
        *reinterpret_cast<uchar4*>(s_l + rr * g0.cols + u0) = o4;
        if (v < r1) *reinterpret_cast<uchar4*>(lab0 + (size_t)v * g0.cols + u0) = o4;
    }
    __syncthreads();
    // adjacency of the band's pixels with their right and lower neighbours
    const int vend = min(r1, g0.rows - 1);
    for (int i = tid; i < (vend - r0) * g0.cols; i += LB_THREADS) {
        int rr, u;
        split_rc(i, g0, rr, u);  // no integer division in the per-pixel loop
        const int v = r0 + rr;
        const float z = s_z[i];
        if (u < g0.cols - 1 && z != 0.f) {
            const int l = s_l[i];
            const int ld = s_l[i + g0.cols], lr = s_l[i + 1];
            if (l != ld && ld != LABEL_NONE) {
                const float zd = s_z[i + g0.cols];
                const float y = (g0.inv_f * (float(v) - g0.disp_v)) * z;
                const float yd = (g0.inv_f * (float(v + 1) - g0.disp_v)) * zd;
                const float disty = sq(z - zd) + sq(y - yd);
                if (disty < prm.conn_dist2_threshold) { atomicOr(&conn[l], 1u << ld); atomicOr(&conn[ld], 1u << l); }
            }
            if (l != lr && lr != LABEL_NONE) {
                const float zr = s_z[i + 1];
                const float x = (g0.inv_f * (float(u) - g0.disp_u)) * z;
                const float xr = (g0.inv_f * (float(u + 1) - g0.disp_u)) * zr;
                const float distx = sq(z - zr) + sq(x - xr);
                if (distx < prm.conn_dist2_threshold) { atomicOr(&conn[l], 1u << lr); atomicOr(&conn[lr], 1u << l); }
            }
        }
    }
    __syncthreads();
    if (tid < NC && conn[tid]) atomicOr(&c.conn[tid], conn[tid]);
}

// labels of the coarser levels (createClustersPyramidUsingKMeans, KMeans.cpp:343-391)
__global__ void __launch_bounds__(256) label_pyr_kernel(Arena a, LevelGeom g) {
    const int pair = blockIdx.y;
    const int frame = a.cur_idx[pair];
    const PairCtl& c = a.ctl[pair];
    __shared__ float cen[NC][3];
    __shared__ float kd[NC][NC];
    if (threadIdx.x < NC)
        for (int r = 0; r < 3; r++) cen[threadIdx.x][r] = c.kmeans[r * NC + threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.x; i < NC * NC; i += blockDim.x) {
        const int la = i / NC, lb = i - la * NC;
        kd[la][lb] = sqnorm3(cen[la][0], cen[la][1], cen[la][2], cen[lb][0], cen[lb][1], cen[lb][2]);
    }
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.P) return;
    const float* depth = a.pyr_d + (size_t)frame * a.pyr_stride + g.off;
    uint8_t* lab = a.labels + (size_t)pair * a.pyr_stride + g.off;
    const float z = __ldg(depth + p);
    uint8_t out = LABEL_NONE;
    if (z != 0.f) {
        int v, u;
        split_rc(p, g, v, u);
        const float x = (g.inv_f * (float(u) - g.disp_u)) * z;
        const float y = (g.inv_f * (float(v) - g.disp_v)) * z;
        int label = 0;
        float min_dist = sqnorm3(cen[0][0], cen[0][1], cen[0][2], z, x, y);
        for (int l = 1; l < NC; l++) {
            if (kd[label][l] > 4.f * min_dist) continue;
            const float dist_here = sqnorm3(cen[l][0], cen[l][1], cen[l][2], z, x, y);
            if (dist_here < min_dist) { label = l; min_dist = dist_here; }
        }
        out = (uint8_t)label;
    }
    lab[p] = out;
}

__global__ void fill_u8_kernel(uint8_t* p, size_t n, uint8_t v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

int launch_kmeans(const Arena& a, const DevParams& p, const LevelGeom* geom, int levels, const LaunchCfg& c) {
    int n = 0;
    if (!p.enable_segmentation) {
        const size_t nb = (size_t)c.n_pairs * a.pyr_stride;
        fill_u8_kernel<<<cdiv(nb, 256), 256, 0, c.stream>>>(a.labels, nb, 0);
        return 1;
    }
    kmeans_kernel<<<c.n_pairs, KM_THREADS, 0, c.stream>>>(a, p, geom[1]); n++;
    {   // rows per band: the band plus its look-ahead row must fit the kernel's shared arrays
        int band = LB_SMEM_PIXELS / geom[0].cols - 1;
        if (band > 24) band = 24;
        const int bands = (geom[0].rows + band - 1) / band;
        label_connect_kernel<<<(unsigned)(bands * c.n_pairs), LB_THREADS, 0, c.stream>>>(a, p, geom[0], geom[1], band, bands); n++;
    }
    for (int l = 2; l < levels; l++) {
        label_pyr_kernel<<<dim3(cdiv(geom[l].P, 256), c.n_pairs), 256, 0, c.stream>>>(a, geom[l]); n++;
    }
    return n;
}
}  // namespace sf
