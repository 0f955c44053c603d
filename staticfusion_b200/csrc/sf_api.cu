// sf_api.cu — C ABI (include/staticfusion_b200.h) and host-side schedule of the B200 StaticFusion solver.
//
// The host enqueues a static schedule (runSolver's loop nest, FrontEnd.cpp:1094-1132, unrolled to its
// maximum trip counts); all data-dependent exits are per-pair device flags, so one batch is solved
// without any host synchronisation.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX: stage ranges for nsys / ncu --nvtx (SF_NVTX=1)

#include "sf_kernels.cuh"

using namespace sf;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) return fail(SF_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

struct sf_ctx {
    sf_params p;
    DevParams dp;
    int device = 0, max_batch = 0, flags = 0;
    int levels = 0;
    LevelGeom geom[MAX_LEVELS];
    cudaStream_t stream = nullptr;
    Arena a{};
    int* d_cur_idx = nullptr;
    int* d_pred_idx = nullptr;
    float* d_twist_in = nullptr;
    uint8_t* d_seed_map = nullptr;
    int n_frames_cap = 0;
    // current batch
    int n_pairs = 0, n_frames = 0;
    bool uploaded = false, solved = false;
    bool is_sequence = false;  // the uploaded batch is a frame sequence (pair k = frames k, k+1)
    int history = 0;           // sf_set_history: run the 5-frame residual stage inside sequence solves
    int stop_step = -1;
    int launches = 0;
    // drop-in trio state
    std::vector<float> h_cur_d, h_cur_i, h_pred_d, h_pred_i;
    bool have_cur = false, have_pred = false, trio_pyr_pred = false;
    float h_twist_old[6] = {0, 0, 0, 0, 0, 0};
    PairOut* h_out = nullptr;  // page-locked staging of the result rows (max_batch entries): device->host copies stay asynchronous
    float* h_pcar = nullptr;   // page-locked staging of perClusterAverageResidual (a copy into pageable memory would block the host)
    struct PendingDownload { bool on = false; int n = 0; float *T = nullptr, *twist = nullptr, *b_segm = nullptr; int *iters = nullptr, *status = nullptr; float* pcar = nullptr; } dl;
    std::vector<int> h_ci, h_pi;  // pair -> frame tables staged for async upload (must outlive the copy)
    // profiling: one event pair per kernel group of the last launch
    bool prof_on = false;
    struct ProfRec { int cls, level; cudaEvent_t e0, e1; };
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    // the static schedule captured once per (batch shape, stop step) and replayed: removes ~500 launch gaps per solve
    struct GraphRec { int n_pairs, n_frames, stop_step, pyramids, history, lanes, prof; cudaGraphExec_t exec; int launches; std::vector<ProfRec> prof_recs; };
    std::vector<GraphRec> graphs;
    bool use_graph = true;
    bool irls_loop = true;      // SF_IRLS_LOOP=0: one launch per pass and iteration instead of the queue-driven loop kernel (A-B measurements)
    int fused_max_tiles = 300;  // levels with at most this many 64-pixel tiles per pair run the fused IRLS kernel (SF_FUSED_MAX_TILES)
    // depth pre-filter scratch (grown on demand)
    uint16_t* d_raw = nullptr;
    float* d_filt = nullptr;
    size_t filt_cap = 0;
    // lanes: a large batch is cut into contiguous pair ranges that run the per-pair schedule on their own streams (forked
    // and joined inside the captured graph), so one range's latency-bound kernels (k-means, solves, pose update, IRLS tail
    // launches) overlap the other ranges' streaming kernels.  Pairs are independent and every sum is an integer sum, so the
    // result does not depend on the cut.  Lane 0 is the context's stream.  SF_LANES overrides the automatic choice.
    static constexpr int MAX_LANES = 4;
    cudaStream_t lane_stream[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
    int* d_lane_gcount = nullptr;    // [MAX_LANES][GCOUNT_CELLS]
    unsigned long long* d_irls_q = nullptr;  // [MAX_LANES][q_cap]
    int* d_iter_list = nullptr;      // [2][max_batch]
    int* d_lane_work_ctr = nullptr;  // [MAX_LANES][MAX_WORK_CTRS]
    int lanes_override = 0, last_lanes = 1;
    // image-sequence loader scratch (grown on demand): raw inputs and converted outputs of sf_convert_frames
    uint8_t* d_cvt = nullptr;
    size_t cvt_cap = 0;
    // copy streams (sf_set_copy_streams): host->device copies of raw frames and device->host copies of results run on their own
    // streams, ordered against the solve stream by events, so that a context's transfers overlap the solves of other contexts
    // (and its own next upload overlaps its current solve) without any host-side wait
    bool copy_streams = false;
    cudaStream_t ul_stream = nullptr, dl_stream = nullptr;
    cudaEvent_t ev_ul = nullptr, ev_cvt = nullptr, ev_solved = nullptr, ev_dl = nullptr;
    bool cvt_recorded = false, dl_recorded = false;
};

static void drop_graphs(sf_ctx* c) {
    for (auto& g : c->graphs) cudaGraphExecDestroy(g.exec);
    c->graphs.clear();
}

static cudaEvent_t prof_event(sf_ctx* c) {
    if (c->ev_used == c->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->ev_pool.push_back(e);
    }
    return c->ev_pool[c->ev_used++];
}
// One event pair per kernel group.  With profiling on, the schedule is still captured and replayed as a CUDA graph (on one
// stream, so that the groups do not overlap): the events are recorded as EXTERNAL event nodes of the graph, i.e. the times
// are those of the product's own execution mode, without the host-side launch gaps plain launches would add to every
// small kernel.
static const char* const kStageNames[SF_PROF_CLASSES] = {"sf:init", "sf:pyramid", "sf:clustering", "sf:warp", "sf:linearise", "sf:irls", "sf:irls_pass2",
                                                         "sf:pose_update", "sf:finish"};
struct ProfScope {
    sf_ctx* c;
    bool on;
    cudaEvent_t e1;
    bool nvtx = false;
    static void record(sf_ctx* c, cudaEvent_t e) {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(c->stream, &st);
        if (st == cudaStreamCaptureStatusActive) cudaEventRecordWithFlags(e, c->stream, cudaEventRecordExternal);
        else cudaEventRecord(e, c->stream);
    }
    ProfScope(sf_ctx* ctx, int cls, int level) : c(ctx), on(ctx->prof_on), e1(nullptr) {
        static const bool want_nvtx = std::getenv("SF_NVTX") != nullptr;  // host-side ranges around the enqueue of every stage
        if (want_nvtx && cls >= 0 && cls < SF_PROF_CLASSES) { nvtxRangePushA(kStageNames[cls]); nvtx = true; }
        if (!on) return;
        cudaEvent_t e0 = prof_event(c);
        e1 = prof_event(c);
        record(c, e0);
        c->prof.push_back({cls, level, e0, e1});
    }
    ~ProfScope() {
        if (on) record(c, e1);
        if (nvtx) nvtxRangePop();
    }
};

static void fill_dev_params(sf_ctx* c) {
    const sf_params& p = c->p;
    DevParams& d = c->dp;
    d.ctf_levels = p.ctf_levels; d.max_iter_per_level = p.max_iter_per_level; d.max_iter_irls = p.max_iter_irls;
    d.use_motion_filter = p.use_motion_filter; d.enable_segmentation = p.enable_segmentation;
    d.k_photometric_res = p.k_photometric_res; d.irls_delta_threshold = p.irls_delta_threshold;
    d.kc_cauchy = p.kc_cauchy; d.kb = p.kb; d.kz = p.kz; d.lambda_reg = p.lambda_reg; d.lambda_prior = p.lambda_prior;
    d.previous_speed_const_weight = p.previous_speed_const_weight; d.previous_speed_eig_weight = p.previous_speed_eig_weight;
    d.outer_exit_threshold = p.outer_exit_threshold;
    for (int i = 0; i < MAX_LEVELS; i++) d.exp_neg_level[i] = expf(-(float)i);  // FrontEnd.cpp:745
    // k-means seeds, KMeans.cpp:76-84 (evaluated at half resolution)
    const int rows_km = p.rows / 2, cols_km = p.cols / 2;
    const unsigned vert_div = (unsigned)std::ceil(std::sqrt((double)NC));
    const float u_div = float(cols_km) / float(NC + 1);
    const float v_div = float(rows_km) / float(vert_div + 1);
    for (unsigned i = 0; i < (unsigned)NC; i++) {
        d.km_u_label[i] = (float)(unsigned)std::round((i + 1) * u_div);
        d.km_v_label[i] = (float)(unsigned)std::round((i % vert_div + 1) * v_div);
    }
    const float t = 0.03f * 120.f / float(p.rows);  // KMeans.cpp:300
    d.conn_dist2_threshold = t * t;
}

static int validate(const sf_params* p, int max_batch) {
    if (!p) return fail(SF_E_INVALID, "params is NULL");
    if (p->rows <= 0 || p->cols <= 0) return fail(SF_E_INVALID, "rows/cols must be positive");
    if (p->ctf_levels < 1 || p->ctf_levels > MAX_LEVELS) return fail(SF_E_INVALID, "ctf_levels must be in [1,8]");
    if (p->enable_segmentation && p->ctf_levels < 2) return fail(SF_E_INVALID, "segmentation needs ctf_levels >= 2 (k-means runs at level 1)");
    if (p->max_iter_per_level < 1 || p->max_iter_irls < 1) return fail(SF_E_INVALID, "iteration counts must be >= 1");
    if (max_batch < 1) return fail(SF_E_INVALID, "max_batch must be >= 1");
    for (int l = 0; l < p->ctf_levels; l++) {
        const int r = p->rows >> l, c = p->cols >> l;
        if (r < 3 || c < 4 || (c % 4) != 0) return fail(SF_E_INVALID, "every pyramid level needs cols % 4 == 0 and rows >= 3");
        if (((p->rows >> l) << l) != p->rows || ((p->cols >> l) << l) != p->cols)
            return fail(SF_E_INVALID, "rows and cols must be divisible by 2^(ctf_levels-1)");
    }
    if ((long long)(p->rows / 2) * (p->rows / 2) + (long long)(p->cols / 2) * (p->cols / 2) >= 1000000LL)
        return fail(SF_E_INVALID, "resolution exceeds the reference's k-means seed range (KMeans.cpp:91)");
    // every IRLS pass launch of one solve takes its own dynamic-item counter (Arena::work_ctr)
    if ((long long)p->ctf_levels * p->max_iter_per_level * (2LL * p->max_iter_irls + 1) > (long long)MAX_WORK_CTRS)
        return fail(SF_E_INVALID, "ctf_levels * max_iter_per_level * (2 * max_iter_irls + 1) exceeds the per-solve launch counters (4096)");
    if (!(p->fovh > 0.f) || !(p->fovh < 3.14f)) return fail(SF_E_INVALID, "fovh must be in (0, pi) radians");
    return SF_OK;
}

// level geometry, constants evaluated exactly as the reference does (FrontEnd.cpp:378-380, 537, 778-780, 874)
static size_t fill_geometry(sf_ctx* c) {
    size_t off = 0;
    const float th = std::tan(0.5f * c->p.fovh);
    for (int l = 0; l < c->levels; l++) {
        LevelGeom& g = c->geom[l];
        g.rows = c->p.rows >> l; g.cols = c->p.cols >> l; g.P = g.rows * g.cols;
        g.f = float(g.cols) / (2.f * th);
        g.inv_f = 2.f * th / float(g.cols);
        g.inv_f_warp = 1.f / g.f;
        g.disp_u = 0.5f * float(g.cols - 1);
        g.disp_v = 0.5f * float(g.rows - 1);
        g.off = off;
        g.cols_magic = (unsigned)((0x100000000ull + (unsigned long long)g.cols - 1ull) / (unsigned long long)g.cols);
        off += (size_t)g.P;
    }
    return off;
}

extern "C" {

int sf_abi_version(void) { return 1; }
const char* sf_last_error(void) { return g_err.c_str(); }

void sf_default_params(sf_params* p, int rows, int cols) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->rows = rows; p->cols = cols;
    int lv = 2, c40 = cols / 40;
    while (c40 > 1) { c40 >>= 1; lv++; }  // log2(cols/40) + 2, FrontEnd.cpp:61
    p->ctf_levels = lv;
    p->max_iter_per_level = 3;   // StaticFusion-datasets.cpp:83
    p->max_iter_irls = 6;        // :89
    p->use_motion_filter = 1;    // :79
    p->enable_segmentation = 1;
    p->fovh = (float)(M_PI * 62.5 / 180.0);  // FrontEnd.cpp:57
    p->k_photometric_res = 0.15f;            // :87
    p->irls_delta_threshold = 0.0015f;       // :88
    p->kc_cauchy = 0.5f; p->kb = 1.5f; p->kz = 1.5f;  // :92-94
    p->lambda_reg = 0.35f; p->lambda_prior = 0.5f;    // :90-91
    p->previous_speed_const_weight = 0.1f;   // :84
    p->previous_speed_eig_weight = 2.f;      // :85
    p->outer_exit_threshold = 0.04f;         // FrontEnd.cpp:1130
}

int sf_create(sf_ctx** out, const sf_params* p, int device, int max_batch, int flags) {
    if (!out) return fail(SF_E_INVALID, "out is NULL");
    *out = nullptr;
    const int rc = validate(p, max_batch);
    if (rc != SF_OK) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(SF_E_CUDA, "no CUDA device: this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(SF_E_INVALID, "device index out of range");
    CU(cudaSetDevice(device));
    sf_ctx* c = new (std::nothrow) sf_ctx();
    if (!c) return fail(SF_E_NOMEM, "host allocation failed");
    c->p = *p; c->device = device; c->max_batch = max_batch; c->flags = flags; c->levels = p->ctf_levels;
    if (const char* e = std::getenv("SF_FUSED_MAX_TILES")) c->fused_max_tiles = std::atoi(e);  // tuning / A-B measurements
    c->fused_max_tiles = std::min(c->fused_max_tiles, 4 * MAX_TILES_PER_WARP_ITEM);
    if (const char* e = std::getenv("SF_IRLS_LOOP")) c->irls_loop = std::atoi(e) != 0;  // the fused kernel's smallest block has 4 warps
    fill_dev_params(c);
    const size_t off = fill_geometry(c);
    Arena& a = c->a;
    a.pyr_stride = off;
    a.P0 = (size_t)c->geom[0].P;
    const int F = max_batch;
    c->n_frames_cap = 2 * F;
    int mb = 1;
    for (int l = 0; l < c->levels; l++) {
        const int it = irls_chunk_iters(c->geom[l].P);
        const int nb = (c->geom[l].P + 1024 * it - 1) / (1024 * it);
        if (nb > mb) mb = nb;
    }
    a.max_blocks = mb;
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return fail(SF_E_CUDA, "cudaGetDeviceProperties failed"); }
        a.num_sms = prop.multiProcessorCount;
    }
    a.trace_steps = p->ctf_levels * p->max_iter_per_level;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return fail(SF_E_CUDA, cudaGetErrorString(e)); }
    auto alloc = [&](void** ptr, size_t bytes) -> bool { return cudaMalloc(ptr, bytes) == cudaSuccess; };
    bool ok = true;
    ok = ok && alloc((void**)&a.pyr_d, sizeof(float) * a.pyr_stride * c->n_frames_cap);
    ok = ok && alloc((void**)&a.pyr_i, sizeof(float) * a.pyr_stride * c->n_frames_cap);
    ok = ok && alloc((void**)&c->d_cur_idx, sizeof(int) * F);
    ok = ok && alloc((void**)&c->d_pred_idx, sizeof(int) * F);
    ok = ok && alloc((void**)&c->d_twist_in, sizeof(float) * 6 * F);
    ok = ok && alloc((void**)&a.labels, a.pyr_stride * F);
    if (c->levels > 1) ok = ok && alloc((void**)&c->d_seed_map, (size_t)c->geom[1].P);
    ok = ok && alloc((void**)&a.acc_d, sizeof(long long) * a.P0 * F);
    ok = ok && alloc((void**)&a.acc_iw, sizeof(unsigned long long) * a.P0 * F);
    ok = ok && alloc((void**)&a.warp_d, sizeof(float) * a.P0 * F);
    ok = ok && alloc((void**)&a.warp_i, sizeof(float) * a.P0 * F);
    ok = ok && alloc((void**)&a.tiles, tiles_per_pair(a.P0) * TILE_BYTES * F);
    ok = ok && alloc((void**)&a.active_list, sizeof(int) * F);
    ok = ok && alloc((void**)&c->d_iter_list, sizeof(int) * 2 * F);
    ok = ok && alloc((void**)&c->d_lane_gcount, sizeof(int) * GCOUNT_CELLS * sf_ctx::MAX_LANES);
    a.q_cap = (int)irls_queue_capacity(F, p->max_iter_irls, a.P0);
    ok = ok && alloc((void**)&c->d_irls_q, sizeof(unsigned long long) * (size_t)a.q_cap * sf_ctx::MAX_LANES);
    ok = ok && alloc((void**)&c->d_lane_work_ctr, sizeof(int) * MAX_WORK_CTRS * sf_ctx::MAX_LANES);
    if (ok && (flags & 1)) ok = alloc((void**)&a.dbg, sizeof(float) * NPLANES * a.P0 * F);
    ok = ok && alloc((void**)&a.ctl, sizeof(PairCtl) * F);
    ok = ok && alloc((void**)&a.out, sizeof(PairOut) * F);
    ok = ok && alloc((void**)&a.b_perpixel, sizeof(float) * a.P0 * F);
    ok = ok && alloc((void**)&a.pcar, sizeof(float) * NC * F);
    ok = ok && alloc((void**)&a.ring_d, sizeof(float) * a.P0 * 5);
    ok = ok && alloc((void**)&a.ring_i, sizeof(float) * a.P0 * 5);
    ok = ok && alloc((void**)&a.ring_T, sizeof(float) * 16 * 5);
    ok = ok && alloc((void**)&a.stepstat, sizeof(int) * 2 * a.trace_steps * F);
    if (ok && (flags & 1)) ok = alloc((void**)&a.trace, sizeof(float) * SF_TRACE_STEP * a.trace_steps * F);
    if (!ok) {
        const std::string msg = std::string("cudaMalloc failed: ") + cudaGetErrorString(cudaGetLastError());
        sf_destroy(c);
        return fail(SF_E_NOMEM, msg);
    }
    a.cur_idx = c->d_cur_idx; a.pred_idx = c->d_pred_idx;
    a.gcount = c->d_lane_gcount; a.work_ctr = c->d_lane_work_ctr; a.irls_q = c->d_irls_q;  // lane 0's slices
    a.iter_list0 = c->d_iter_list; a.iter_list1 = c->d_iter_list + F;
    if (const char* e = std::getenv("SF_LANES")) c->lanes_override = std::atoi(e);
    c->lane_stream[0] = c->stream;
    for (int l = 1; l < sf_ctx::MAX_LANES; l++) {
        if (cudaStreamCreateWithFlags(&c->lane_stream[l], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_join[l], cudaEventDisableTiming) != cudaSuccess) { sf_destroy(c); return fail(SF_E_CUDA, "lane stream creation failed"); }
    }
    if (cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess) { sf_destroy(c); return fail(SF_E_CUDA, "event creation failed"); }
    if (c->levels > 1) {  // seed labelling of level 1, KMeans.cpp:87-101 (integer arithmetic; 24 = no seed within range)
        const LevelGeom& g1 = c->geom[1];
        std::vector<uint8_t> seed((size_t)g1.P);
        for (int v = 0; v < g1.rows; v++)
            for (int u = 0; u < g1.cols; u++) {
                unsigned min_dist = 1000000u;
                int lab = NC;
                for (int l = 0; l < NC; l++) {
                    const int dv = v - (int)c->dp.km_v_label[l], du = u - (int)c->dp.km_u_label[l];
                    const unsigned qd = (unsigned)(dv * dv + du * du);
                    if (qd < min_dist) { lab = l; min_dist = qd; }
                }
                seed[(size_t)v * g1.cols + u] = (uint8_t)lab;
            }
        cudaMemcpyAsync(c->d_seed_map, seed.data(), seed.size(), cudaMemcpyHostToDevice, c->stream);
        cudaStreamSynchronize(c->stream);  // the staging vector goes out of scope
        a.seed_map = c->d_seed_map;
    }
    // splat accumulators are kept zero between uses (warp_normalise clears what it reads)
    cudaMemsetAsync(a.acc_d, 0, sizeof(long long) * a.P0 * F, c->stream);
    cudaMemsetAsync(a.acc_iw, 0, sizeof(unsigned long long) * a.P0 * F, c->stream);
    cudaMemsetAsync(a.tiles, 0xff, tiles_per_pair(a.P0) * TILE_BYTES * F, c->stream);  // every label byte = invalid
    cudaMemsetAsync(c->d_lane_gcount, 0, sizeof(int) * GCOUNT_CELLS * sf_ctx::MAX_LANES, c->stream);
    cudaMemsetAsync(c->d_irls_q, 0, sizeof(unsigned long long) * (size_t)a.q_cap * sf_ctx::MAX_LANES, c->stream);  // generation 0 = never valid
    cudaMemsetAsync(a.pcar, 0xff, sizeof(float) * NC * F, c->stream);  // all-ones = quiet NaN (FrontEnd.cpp:105)
    cudaMemsetAsync(a.ring_d, 0, sizeof(float) * a.P0 * 5, c->stream);
    cudaMemsetAsync(a.ring_i, 0, sizeof(float) * a.P0 * 5, c->stream);
    cudaMemsetAsync(a.ring_T, 0, sizeof(float) * 16 * 5, c->stream);
    if (a.dbg) cudaMemsetAsync(a.dbg, 0, sizeof(float) * NPLANES * a.P0 * F, c->stream);
    e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { sf_destroy(c); return fail(SF_E_CUDA, cudaGetErrorString(e)); }
    sf::prepare_kernels();  // function attributes are per device: set them for this context's device (cudaSetDevice above)
    if (cudaHostAlloc((void**)&c->h_out, sizeof(PairOut) * F, cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void**)&c->h_pcar, sizeof(float) * NC * F, cudaHostAllocDefault) != cudaSuccess) { sf_destroy(c); return fail(SF_E_NOMEM, "cudaHostAlloc failed"); }
    *out = c;
    return SF_OK;
}

void sf_destroy(sf_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    Arena& a = c->a;
    cudaFree(a.pyr_d); cudaFree(a.pyr_i); cudaFree(c->d_cur_idx); cudaFree(c->d_pred_idx); cudaFree(c->d_twist_in); cudaFree(c->d_seed_map);
    cudaFree(a.labels); cudaFree(a.acc_d); cudaFree(a.acc_iw); cudaFree(a.warp_d); cudaFree(a.warp_i); cudaFree(a.tiles); cudaFree(a.dbg); cudaFree(a.active_list);
    cudaFree(a.ctl); cudaFree(a.out); cudaFree(a.b_perpixel); cudaFree(a.pcar); cudaFree(a.ring_d); cudaFree(a.ring_i); cudaFree(a.ring_T);
    cudaFree(a.trace); cudaFree(a.stepstat); cudaFree(c->d_raw); cudaFree(c->d_filt); cudaFree(c->d_cvt);
    drop_graphs(c);
    cudaFree(c->d_lane_gcount); cudaFree(c->d_lane_work_ctr); cudaFree(c->d_iter_list); cudaFree(c->d_irls_q);
    if (c->h_out) cudaFreeHost(c->h_out);
    if (c->h_pcar) cudaFreeHost(c->h_pcar);
    for (int l = 1; l < sf_ctx::MAX_LANES; l++) { if (c->lane_stream[l]) cudaStreamDestroy(c->lane_stream[l]); if (c->ev_join[l]) cudaEventDestroy(c->ev_join[l]); }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    if (c->ul_stream) cudaStreamDestroy(c->ul_stream);
    if (c->dl_stream) cudaStreamDestroy(c->dl_stream);
    for (cudaEvent_t e : {c->ev_ul, c->ev_cvt, c->ev_solved, c->ev_dl}) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int sf_set_params(sf_ctx* c, const sf_params* p) {
    if (!c || !p) return fail(SF_E_INVALID, "NULL argument");
    if (p->rows != c->p.rows || p->cols != c->p.cols || p->ctf_levels != c->p.ctf_levels ||
        p->max_iter_per_level != c->p.max_iter_per_level)
        return fail(SF_E_INVALID, "rows, cols, ctf_levels and max_iter_per_level are fixed at sf_create");
    if (std::memcmp(p, &c->p, sizeof(sf_params)) == 0) return SF_OK;  // nothing changed: the captured graphs stay valid
    const int rc = validate(p, c->max_batch);
    if (rc != SF_OK) return rc;
    c->p = *p;
    fill_dev_params(c);
    fill_geometry(c);  // fovh may have changed: focal lengths of every level
    drop_graphs(c);    // kernel arguments are baked into captured graphs
    return SF_OK;
}

// ---------------------------------------------------------------------------------------------
// upload
// ---------------------------------------------------------------------------------------------
static int upload_stack(sf_ctx* c, float* dst_pyr, int first_frame, int frame_step, int n, const float* src, int in_space) {
    // image k -> level-0 slot of frame first_frame + k*frame_step
    const size_t w = sizeof(float) * c->a.P0;
    const cudaMemcpyKind kind = (in_space == SF_MEM_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    CU(cudaMemcpy2DAsync(dst_pyr + (size_t)first_frame * c->a.pyr_stride, sizeof(float) * c->a.pyr_stride * frame_step, src, w, w,
                         (size_t)n, kind, c->stream));
    return SF_OK;
}

static int upload_twist(sf_ctx* c, int n_pairs, const float* twist_old_in) {
    if (twist_old_in) CU(cudaMemcpyAsync(c->d_twist_in, twist_old_in, sizeof(float) * 6 * n_pairs, cudaMemcpyHostToDevice, c->stream));
    else CU(cudaMemsetAsync(c->d_twist_in, 0, sizeof(float) * 6 * n_pairs, c->stream));
    return SF_OK;
}

int sf_upload_pairs(sf_ctx* c, int n_pairs, const float* depth_cur, const float* inten_cur, const float* depth_pred,
                    const float* inten_pred, int in_space, const float* twist_old_in) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    if (n_pairs < 1 || n_pairs > c->max_batch) return fail(SF_E_INVALID, "n_pairs must be in [1, max_batch]");
    if (!depth_cur || !inten_cur || !depth_pred || !inten_pred) return fail(SF_E_INVALID, "NULL image pointer");
    CU(cudaSetDevice(c->device));
    // frames: current of pair k = 2k, prediction = 2k+1
    int rc;
    if ((rc = upload_stack(c, c->a.pyr_d, 0, 2, n_pairs, depth_cur, in_space))) return rc;
    if ((rc = upload_stack(c, c->a.pyr_i, 0, 2, n_pairs, inten_cur, in_space))) return rc;
    if ((rc = upload_stack(c, c->a.pyr_d, 1, 2, n_pairs, depth_pred, in_space))) return rc;
    if ((rc = upload_stack(c, c->a.pyr_i, 1, 2, n_pairs, inten_pred, in_space))) return rc;
    std::vector<int>&ci = c->h_ci, &pi = c->h_pi;
    if ((int)ci.size() != n_pairs || ci[0] != 0 || pi[0] != 1) {
        CU(cudaStreamSynchronize(c->stream));  // a previous async copy may still read the tables
        ci.resize(n_pairs); pi.resize(n_pairs);
        for (int k = 0; k < n_pairs; k++) { ci[k] = 2 * k; pi[k] = 2 * k + 1; }
    }
    CU(cudaMemcpyAsync(c->d_cur_idx, ci.data(), sizeof(int) * n_pairs, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_pred_idx, pi.data(), sizeof(int) * n_pairs, cudaMemcpyHostToDevice, c->stream));
    if ((rc = upload_twist(c, n_pairs, twist_old_in))) return rc;
    if (twist_old_in) CU(cudaStreamSynchronize(c->stream));  // caller may free twist_old_in on return
    c->n_pairs = n_pairs; c->n_frames = 2 * n_pairs; c->uploaded = true; c->solved = false; c->is_sequence = false;
    return SF_OK;
}

int sf_upload_sequence(sf_ctx* c, int n_frames, const float* depth, const float* inten, int in_space, const float* twist_old_in) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    if (n_frames < 2 || n_frames - 1 > c->max_batch) return fail(SF_E_INVALID, "n_frames-1 must be in [1, max_batch]");
    if (!depth || !inten) return fail(SF_E_INVALID, "NULL image pointer");
    CU(cudaSetDevice(c->device));
    int rc;
    if ((rc = upload_stack(c, c->a.pyr_d, 0, 1, n_frames, depth, in_space))) return rc;
    if ((rc = upload_stack(c, c->a.pyr_i, 0, 1, n_frames, inten, in_space))) return rc;
    const int n_pairs = n_frames - 1;
    std::vector<int>&ci = c->h_ci, &pi = c->h_pi;
    if ((int)ci.size() != n_pairs || ci[0] != 1 || pi[0] != 0) {
        CU(cudaStreamSynchronize(c->stream));
        ci.resize(n_pairs); pi.resize(n_pairs);
        for (int k = 0; k < n_pairs; k++) { ci[k] = k + 1; pi[k] = k; }  // prediction := previous raw frame
    }
    CU(cudaMemcpyAsync(c->d_cur_idx, ci.data(), sizeof(int) * n_pairs, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_pred_idx, pi.data(), sizeof(int) * n_pairs, cudaMemcpyHostToDevice, c->stream));
    if ((rc = upload_twist(c, n_pairs, twist_old_in))) return rc;
    if (twist_old_in) CU(cudaStreamSynchronize(c->stream));
    c->n_pairs = n_pairs; c->n_frames = n_frames; c->uploaded = true; c->solved = false; c->is_sequence = true;
    return SF_OK;
}

// ---------------------------------------------------------------------------------------------
// schedule
// ---------------------------------------------------------------------------------------------
// the arena as one lane sees it: every per-pair array starts at the lane's first pair, control cells are the lane's own
static Arena lane_arena(const sf_ctx* c, int lane, int lo) {
    Arena a = c->a;
    const size_t o = (size_t)lo;
    a.cur_idx += o; a.pred_idx += o;
    a.labels += o * a.pyr_stride;
    a.acc_d += o * a.P0; a.acc_iw += o * a.P0; a.warp_d += o * a.P0; a.warp_i += o * a.P0;
    a.tiles += o * tiles_per_pair(a.P0) * TILE_BYTES;
    if (a.dbg) a.dbg += o * NPLANES * a.P0;
    a.active_list += o; a.iter_list0 += o; a.iter_list1 += o;
    a.gcount = c->d_lane_gcount + GCOUNT_CELLS * lane;
    a.irls_q = c->d_irls_q + (size_t)a.q_cap * lane;
    a.work_ctr = c->d_lane_work_ctr + (size_t)MAX_WORK_CTRS * lane;
    a.ctl += o; a.out += o;
    a.b_perpixel += o * a.P0;
    a.pcar += o * NC;
    if (a.trace) a.trace += o * SF_TRACE_STEP * a.trace_steps;
    a.stepstat += o * 2 * a.trace_steps;
    return a;
}

static int choose_lanes(const sf_ctx* c) {
    if (c->prof_on) return 1;  // per-kernel events time each kernel alone on one stream
    // With the queue-driven IRLS loop kernel a lane's kernels are self-contained (no trail of tail launches to overlap): one
    // lane with full-size launches is fastest and callers overlap whole batches on several contexts instead (measured, 512
    // QVGA pairs, 2 contexts: 1 / 2 / 3 / 4 lanes = 5.10 / 5.23 / 5.35 / 5.71 ms per step).  The per-launch passes keep the old rule.
    int n = c->lanes_override > 0 ? c->lanes_override : (c->irls_loop ? 1 : (c->n_pairs >= 384 ? 3 : c->n_pairs >= 192 ? 2 : 1));
    if (n > sf_ctx::MAX_LANES) n = sf_ctx::MAX_LANES;
    if (n > c->n_pairs) n = c->n_pairs;
    return n < 1 ? 1 : n;
}

// the per-pair part of runSolver for the pairs [lo, lo + n) on `stream`; returns false when the debug stop step cut it short
static bool enqueue_lane(sf_ctx* c, const Arena& a, cudaStream_t stream, int lo, int n_lane, int* n_out) {
    int ctr = 0;
    const LaunchCfg cfg{stream, n_lane, c->n_frames, &ctr};
    const DevParams& dp = c->dp;
    int n = 0;
    { ProfScope ps(c, 0, 0); n += launch_init_pairs(a, dp, c->d_twist_in + 6 * (size_t)lo, cfg); }
    { ProfScope ps(c, 2, 0); n += launch_kmeans(a, dp, c->geom, c->levels, cfg); }
    bool stop = false;
    for (int i = 0; i < c->levels && !stop; i++)
        for (int k = 0; k < c->p.max_iter_per_level && !stop; k++) {
            const int image_level = c->levels - i - 1;  // FrontEnd.cpp:1100
            const LevelGeom& g = c->geom[image_level];
            const int first = (i == 0 && k == 0) ? 1 : 0;
            if (!first) { ProfScope ps(c, 3, image_level); n += launch_step_begin(a, i, k, cfg); n += launch_warp(a, g, cfg); }
            else n += launch_step_begin(a, i, k, cfg);
            { ProfScope ps(c, 4, image_level); n += launch_linearise(a, dp, g, first, cfg); n += launch_step_prep(a, dp, i, k, cfg); }
            if (i * c->p.max_iter_per_level + k == c->stop_step) { stop = true; break; }
            if ((int)tiles_per_pair((size_t)g.P) <= c->fused_max_tiles) {  // small level: one block runs the pair's whole IRLS loop
                ProfScope ps(c, 5, image_level);
                n += launch_irls_fused(a, dp, g, i, k, cfg);
            } else if (c->irls_loop) {  // large level: ONE launch runs the IRLS loops of all pairs through a device-side item queue
                ProfScope ps(c, 5, image_level);
                n += launch_irls_loop(a, dp, g, i, k, cfg);
            } else
                for (int it = 1; it <= c->p.max_iter_irls; it++) {
                    { ProfScope ps(c, 5, image_level); n += launch_irls_pass1(a, dp, g, i, k, it, cfg); }
                    { ProfScope ps(c, 6, image_level); n += launch_irls_pass2(a, dp, g, i, k, it, cfg); }
                }
            { ProfScope ps(c, 7, image_level); n += launch_pose_update(a, dp, i, k, cfg); }
        }
    { ProfScope ps(c, 8, 0); n += launch_finish(a, dp, c->geom[0], cfg); }
    *n_out += n;
    return !stop;
}

static int enqueue_solve(sf_ctx* c, bool build_pyramids) {
    int ctr = 0;
    const LaunchCfg cfg{c->stream, c->n_pairs, c->n_frames, &ctr};
    const Arena& a = c->a;
    const DevParams& dp = c->dp;
    int n = 0;
    c->prof.clear();
    c->ev_used = 0;
    if (build_pyramids) { ProfScope ps(c, 1, 0); n += launch_pyramids(a, c->geom, c->levels, cfg); }
    const int lanes = choose_lanes(c);
    c->last_lanes = lanes;
    bool complete = true;
    if (lanes == 1) complete = enqueue_lane(c, a, c->stream, 0, c->n_pairs, &n);
    else {
        CU(cudaEventRecord(c->ev_fork, c->stream));  // the pyramids are ready
        const int per = (c->n_pairs + lanes - 1) / lanes;
        for (int l = 0; l < lanes; l++) {
            const int lo = l * per, nl = std::min(per, c->n_pairs - lo);
            if (nl <= 0) continue;
            if (l) CU(cudaStreamWaitEvent(c->lane_stream[l], c->ev_fork, 0));
            complete = enqueue_lane(c, lane_arena(c, l, lo), c->lane_stream[l], lo, nl, &n) && complete;
            if (l) { CU(cudaEventRecord(c->ev_join[l], c->lane_stream[l])); CU(cudaStreamWaitEvent(c->stream, c->ev_join[l], 0)); }
        }
    }
    {
        ProfScope ps(c, 8, 0);
        if (c->history && c->is_sequence && complete) n += launch_history(a, dp, c->geom[0], 0, 0, cfg);  // StaticFusion-datasets.cpp:175-177
        n += launch_segm_image(a, c->geom[0], cfg);
    }
    c->launches = n;
    CU(cudaGetLastError());
    return SF_OK;
}

static int launch_solve(sf_ctx* c, bool build_pyramids) {
    static const bool prof_plain = std::getenv("SF_PROF_PLAIN") != nullptr;  // A-B: per-kernel events around plain launches
    if (!c->use_graph || (c->prof_on && prof_plain)) return enqueue_solve(c, build_pyramids);
    for (const auto& g : c->graphs)
        if (g.n_pairs == c->n_pairs && g.n_frames == c->n_frames && g.stop_step == c->stop_step && g.pyramids == (int)build_pyramids &&
            g.history == (c->history && c->is_sequence) && g.lanes == choose_lanes(c) && g.prof == (int)c->prof_on) {
            CU(cudaGraphLaunch(g.exec, c->stream));
            c->launches = g.launches;
            if (c->prof_on) c->prof = g.prof_recs;  // the graph's external event nodes: re-recorded by this replay
            return SF_OK;
        }
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue_solve(c, build_pyramids);
    const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    if (rc != SF_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail(SF_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    cudaGraphExec_t exec = nullptr;
    const cudaError_t e2 = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e2 != cudaSuccess) return fail(SF_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e2));
    c->graphs.push_back({c->n_pairs, c->n_frames, c->stop_step, (int)build_pyramids, (int)(c->history && c->is_sequence), choose_lanes(c), (int)c->prof_on, exec,
                         c->launches, c->prof_on ? c->prof : std::vector<sf_ctx::ProfRec>()});
    CU(cudaGraphLaunch(exec, c->stream));
    return SF_OK;
}

int sf_launch(sf_ctx* c) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    if (!c->uploaded) return fail(SF_E_STATE, "no batch uploaded");
    CU(cudaSetDevice(c->device));
    // batched solves carry no history between calls: perClusterAverageResidual starts as NaN (FrontEnd.cpp:105)
    if (c->copy_streams && c->dl_recorded) CU(cudaStreamWaitEvent(c->stream, c->ev_dl, 0));  // the last results have left the arena
    CU(cudaMemsetAsync(c->a.pcar, 0xff, sizeof(float) * NC * c->n_pairs, c->stream));
    const int rc = launch_solve(c, true);
    if (rc == SF_OK) {
        c->solved = true;
        if (c->copy_streams) CU(cudaEventRecord(c->ev_solved, c->stream));
    }
    return rc;
}

int sf_set_copy_streams(sf_ctx* c, int on) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    if (on && !c->ul_stream) {
        CU(cudaStreamCreateWithFlags(&c->ul_stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&c->dl_stream, cudaStreamNonBlocking));
        for (cudaEvent_t* e : {&c->ev_ul, &c->ev_cvt, &c->ev_solved, &c->ev_dl}) CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    if (c->dl_stream) CU(cudaStreamSynchronize(c->dl_stream));
    if (c->ul_stream) CU(cudaStreamSynchronize(c->ul_stream));
    c->copy_streams = on != 0;
    c->cvt_recorded = c->dl_recorded = false;
    return SF_OK;
}

int sf_set_history(sf_ctx* c, int on) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    c->history = on ? 1 : 0;
    return SF_OK;
}

int sf_get_per_cluster_average_residual(sf_ctx* c, float* out) {
    if (!c || !out) return fail(SF_E_INVALID, "NULL argument");
    if (!c->solved) return fail(SF_E_STATE, "nothing has been solved");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(out, c->a.pcar, sizeof(float) * NC * c->n_pairs, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SF_OK;
}

int sf_sync(sf_ctx* c) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    CU(cudaStreamSynchronize(c->stream));
    return SF_OK;
}

uint64_t sf_stream(sf_ctx* c) { return c ? (uint64_t)(uintptr_t)c->stream : 0; }
int sf_last_launch_count(sf_ctx* c) { return c ? c->launches : 0; }
int sf_last_lane_count(sf_ctx* c) { return c ? choose_lanes(c) : 0; }

// split phase: _begin enqueues every device->host copy behind the solve (no host wait), _end waits and unpacks the rows
int sf_download_range_begin(sf_ctx* c, int first_pair, int n, float* T_odometry, float* twist_old_out, float* b_segm, float* b_perpixel,
                            uint8_t* labels_u8, int out_space, int* irls_iters, int* status, float* per_cluster_residual) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    if (!c->solved) return fail(SF_E_STATE, "nothing has been solved");
    if (c->dl.on) return fail(SF_E_STATE, "a download is already in flight: call sf_download_range_end first");
    if (first_pair < 0 || n < 0 || first_pair + n > c->n_pairs) return fail(SF_E_INVALID, "pair range outside the last solve");
    c->dl = {true, n, T_odometry, twist_old_out, b_segm, irls_iters, status, per_cluster_residual};
    if (n == 0) return SF_OK;
    CU(cudaSetDevice(c->device));
    const Arena& a = c->a;
    cudaStream_t st = c->stream;
    if (c->copy_streams) {  // the copies run on the download stream, behind the solve
        CU(cudaStreamWaitEvent(c->dl_stream, c->ev_solved, 0));
        st = c->dl_stream;
    }
    CU(cudaMemcpyAsync(c->h_out, a.out + first_pair, sizeof(PairOut) * n, cudaMemcpyDeviceToHost, st));
    const cudaMemcpyKind kind = (out_space == SF_MEM_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (b_perpixel) CU(cudaMemcpyAsync(b_perpixel, a.b_perpixel + (size_t)first_pair * a.P0, sizeof(float) * a.P0 * n, kind, st));
    if (labels_u8) CU(cudaMemcpy2DAsync(labels_u8, a.P0, a.labels + (size_t)first_pair * a.pyr_stride, a.pyr_stride, a.P0, (size_t)n, kind, st));
    if (per_cluster_residual) CU(cudaMemcpyAsync(c->h_pcar, a.pcar + (size_t)first_pair * NC, sizeof(float) * NC * n, cudaMemcpyDeviceToHost, st));
    if (c->copy_streams) { CU(cudaEventRecord(c->ev_dl, c->dl_stream)); c->dl_recorded = true; }
    return SF_OK;
}

int sf_download_range_end(sf_ctx* c) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    if (!c->dl.on) return fail(SF_E_STATE, "no download in flight");
    const sf_ctx::PendingDownload d = c->dl;
    c->dl.on = false;
    if (c->copy_streams && d.n) CU(cudaStreamSynchronize(c->dl_stream));
    else CU(cudaStreamSynchronize(c->stream));
    if (d.pcar && d.n) std::memcpy(d.pcar, c->h_pcar, sizeof(float) * NC * d.n);
    for (int k = 0; k < d.n; k++) {
        const PairOut& o = c->h_out[k];
        if (d.T)
            for (int r = 0; r < 4; r++)
                for (int q = 0; q < 4; q++) d.T[k * 16 + q * 4 + r] = o.T[r * 4 + q];  // row-major -> Eigen column-major
        if (d.twist) std::memcpy(d.twist + k * 6, o.twist_old, sizeof(float) * 6);
        if (d.b_segm) std::memcpy(d.b_segm + k * NC, o.b_segm, sizeof(float) * NC);
        if (d.iters) d.iters[k] = o.irls_iters;
        if (d.status) d.status[k] = o.status;
    }
    return SF_OK;
}

int sf_download_range(sf_ctx* c, int first_pair, int n, float* T_odometry, float* twist_old_out, float* b_segm, float* b_perpixel,
                      uint8_t* labels_u8, int out_space, int* irls_iters, int* status, float* per_cluster_residual) {
    const int rc = sf_download_range_begin(c, first_pair, n, T_odometry, twist_old_out, b_segm, b_perpixel, labels_u8, out_space, irls_iters,
                                           status, per_cluster_residual);
    if (rc) return rc;
    return sf_download_range_end(c);
}

int sf_download(sf_ctx* c, float* T_odometry, float* twist_old_out, float* b_segm, float* b_perpixel, uint8_t* labels_u8,
                int out_space, int* irls_iters, int* status) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    return sf_download_range(c, 0, c->n_pairs, T_odometry, twist_old_out, b_segm, b_perpixel, labels_u8, out_space, irls_iters, status, nullptr);
}

int sf_solve_batch(sf_ctx* c, int n_pairs, const float* depth_cur, const float* inten_cur, const float* depth_pred,
                   const float* inten_pred, int in_space, const float* twist_old_in, float* T_odometry, float* twist_old_out,
                   float* b_segm, float* b_perpixel, uint8_t* labels_u8, int out_space, int* irls_iters, int* status) {
    int rc = sf_upload_pairs(c, n_pairs, depth_cur, inten_cur, depth_pred, inten_pred, in_space, twist_old_in);
    if (rc) return rc;
    if ((rc = sf_launch(c))) return rc;
    return sf_download(c, T_odometry, twist_old_out, b_segm, b_perpixel, labels_u8, out_space, irls_iters, status);
}

int sf_solve_sequence(sf_ctx* c, int n_frames, const float* depth, const float* inten, int in_space, const float* twist_old_in,
                      float* T_odometry, float* twist_old_out, float* b_segm, float* b_perpixel, uint8_t* labels_u8,
                      int out_space, int* irls_iters, int* status) {
    int rc = sf_upload_sequence(c, n_frames, depth, inten, in_space, twist_old_in);
    if (rc) return rc;
    if ((rc = sf_launch(c))) return rc;
    return sf_download(c, T_odometry, twist_old_out, b_segm, b_perpixel, labels_u8, out_space, irls_iters, status);
}

// ---------------------------------------------------------------------------------------------
// drop-in trio (one pair, host buffers, reference call order)
// ---------------------------------------------------------------------------------------------
static void to_row_major(const float* src, float* dst, int rows, int cols, int col_major) {
    if (!col_major) { std::memcpy(dst, src, sizeof(float) * rows * cols); return; }
    for (int v = 0; v < rows; v++)
        for (int u = 0; u < cols; u++) dst[(size_t)v * cols + u] = src[(size_t)u * rows + v];
}

int sf_set_current(sf_ctx* c, const float* depth, const float* intensity, int col_major) {
    if (!c || !depth || !intensity) return fail(SF_E_INVALID, "NULL argument");
    const size_t n = c->a.P0;
    c->h_cur_d.resize(n); c->h_cur_i.resize(n);
    to_row_major(depth, c->h_cur_d.data(), c->p.rows, c->p.cols, col_major);
    to_row_major(intensity, c->h_cur_i.data(), c->p.rows, c->p.cols, col_major);
    c->have_cur = true;
    return SF_OK;
}

int sf_set_prediction(sf_ctx* c, const float* depth, const float* intensity, int col_major) {
    if (!c || !depth || !intensity) return fail(SF_E_INVALID, "NULL argument");
    const size_t n = c->a.P0;
    c->h_pred_d.resize(n); c->h_pred_i.resize(n);
    to_row_major(depth, c->h_pred_d.data(), c->p.rows, c->p.cols, col_major);
    to_row_major(intensity, c->h_pred_i.data(), c->p.rows, c->p.cols, col_major);
    c->have_pred = true;
    c->trio_pyr_pred = false;
    return SF_OK;
}

int sf_set_twist_old(sf_ctx* c, const float t[6]) {
    if (!c || !t) return fail(SF_E_INVALID, "NULL argument");
    std::memcpy(c->h_twist_old, t, sizeof(float) * 6);
    return SF_OK;
}

int sf_create_image_pyramid(sf_ctx* c, int old_im) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    CU(cudaSetDevice(c->device));
    // frame slot 0 = current, 1 = prediction (pair 0)
    const int frame = old_im ? 1 : 0;
    if (old_im ? !c->have_pred : !c->have_cur) return fail(SF_E_STATE, "input images were not set");
    const float* d = old_im ? c->h_pred_d.data() : c->h_cur_d.data();
    const float* i = old_im ? c->h_pred_i.data() : c->h_cur_i.data();
    int rc;
    if ((rc = upload_stack(c, c->a.pyr_d, frame, 1, 1, d, SF_MEM_HOST))) return rc;
    if ((rc = upload_stack(c, c->a.pyr_i, frame, 1, 1, i, SF_MEM_HOST))) return rc;
    // build this frame's pyramid only: temporarily view the arena from `frame`
    Arena a = c->a;
    a.pyr_d += (size_t)frame * a.pyr_stride;
    a.pyr_i += (size_t)frame * a.pyr_stride;
    const LaunchCfg cfg{c->stream, 1, 1, nullptr};
    launch_pyramids(a, c->geom, c->levels, cfg);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    if (old_im) c->trio_pyr_pred = true;
    return SF_OK;
}

int sf_run_solver(sf_ctx* c, int create_image_pyr) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    if (!c->have_cur || !c->have_pred) return fail(SF_E_STATE, "input images were not set");
    if (!c->trio_pyr_pred) return fail(SF_E_STATE, "createImagePyramid(true) must run before runSolver (StaticFusion-datasets.cpp:171-173)");
    CU(cudaSetDevice(c->device));
    int rc;
    if (create_image_pyr) {
        if ((rc = sf_create_image_pyramid(c, 0))) return rc;
    }
    const int ci = 0, pi = 1;
    CU(cudaMemcpyAsync(c->d_cur_idx, &ci, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_pred_idx, &pi, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_twist_in, c->h_twist_old, sizeof(float) * 6, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->n_pairs = 1; c->n_frames = 2; c->uploaded = true; c->is_sequence = false;
    rc = launch_solve(c, false);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    c->solved = true;
    return SF_OK;
}

int sf_build_segm_image(sf_ctx* c) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    if (!c->solved) return fail(SF_E_STATE, "runSolver has not run");
    CU(cudaSetDevice(c->device));
    // every solve ends with this kernel already; re-run it so that a computeResidualsAgainstPreviousImage call made
    // in between (StaticFusion-datasets.cpp:175-180) takes effect
    const LaunchCfg cfg{c->stream, c->n_pairs, c->n_frames, nullptr};
    launch_segm_image(c->a, c->geom[0], cfg);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return SF_OK;
}

int sf_buffer_set(sf_ctx* c, int slot, const float* depth, const float* intensity, const float T[16], int col_major) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    if ((depth == nullptr) != (intensity == nullptr)) return fail(SF_E_INVALID, "depth and intensity must be given together");
    CU(cudaSetDevice(c->device));
    const int b = ((slot % 5) + 5) % 5;
    const size_t n = c->a.P0;
    std::vector<float> d, i;
    if (depth) {
        d.resize(n); i.resize(n);
        to_row_major(depth, d.data(), c->p.rows, c->p.cols, col_major);
        to_row_major(intensity, i.data(), c->p.rows, c->p.cols, col_major);
        CU(cudaMemcpyAsync(c->a.ring_d + (size_t)b * n, d.data(), sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->a.ring_i + (size_t)b * n, i.data(), sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    }
    float Tr[16];
    for (int r = 0; r < 4; r++)
        for (int q = 0; q < 4; q++) Tr[r * 4 + q] = T ? T[q * 4 + r] : (r == q ? 1.f : 0.f);  // Eigen column-major -> row-major
    CU(cudaMemcpyAsync(c->a.ring_T + 16 * b, Tr, sizeof(float) * 16, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SF_OK;
}

int sf_filter_depth(sf_ctx* c, int n_images, const uint16_t* depth_mm, int in_space, float max_depth_m, float* depth_out, int out_space,
                    int col_major_out) {
    if (!c || !depth_mm || !depth_out) return fail(SF_E_INVALID, "NULL argument");
    if (n_images < 1) return fail(SF_E_INVALID, "n_images must be >= 1");
    if (col_major_out && out_space == SF_MEM_DEVICE) return fail(SF_E_INVALID, "column-major output is a host-side conversion");
    CU(cudaSetDevice(c->device));
    const size_t P = c->a.P0, n = (size_t)n_images;
    const bool need_in = in_space != SF_MEM_DEVICE, need_out = out_space != SF_MEM_DEVICE;
    if ((need_in || need_out) && c->filt_cap < n) {
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_raw); cudaFree(c->d_filt); c->d_raw = nullptr; c->d_filt = nullptr; c->filt_cap = 0;
        CU(cudaMalloc((void**)&c->d_raw, sizeof(uint16_t) * P * n));
        CU(cudaMalloc((void**)&c->d_filt, sizeof(float) * P * n));
        c->filt_cap = n;
    }
    const uint16_t* src = depth_mm;
    if (need_in) { CU(cudaMemcpyAsync(c->d_raw, depth_mm, sizeof(uint16_t) * P * n, cudaMemcpyHostToDevice, c->stream)); src = c->d_raw; }
    float* dst = need_out ? c->d_filt : depth_out;
    launch_filter_depth(src, dst, c->p.rows, c->p.cols, n_images, P, P, max_depth_m, c->stream);
    CU(cudaGetLastError());
    if (need_out) {
        if (!col_major_out) CU(cudaMemcpyAsync(depth_out, c->d_filt, sizeof(float) * P * n, cudaMemcpyDeviceToHost, c->stream));
        else {
            std::vector<float> tmp(P * n);
            CU(cudaMemcpyAsync(tmp.data(), c->d_filt, sizeof(float) * P * n, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            const int rows = c->p.rows, cols = c->p.cols;
            for (size_t k = 0; k < n; k++)
                for (int v = 0; v < rows; v++)
                    for (int u = 0; u < cols; u++) depth_out[k * P + (size_t)u * rows + v] = tmp[k * P + (size_t)v * cols + u];
        }
    }
    CU(cudaStreamSynchronize(c->stream));
    return SF_OK;
}

// ---------------------------------------------------------------------------------------------
// image-sequence loader (FrontEnd.cpp:216-254)
// ---------------------------------------------------------------------------------------------
static int cvt_reserve(sf_ctx* c, size_t bytes) {
    if (c->cvt_cap >= bytes) return SF_OK;
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_cvt); c->d_cvt = nullptr; c->cvt_cap = 0;
    CU(cudaMalloc((void**)&c->d_cvt, bytes));
    c->cvt_cap = bytes;
    return SF_OK;
}
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int sf_convert_frames(sf_ctx* c, int n_images, const uint8_t* bgr, const uint16_t* depth_raw, int res_factor, int in_space, float* intensity,
                      float* depth, uint16_t* depth_mm, uint8_t* color_full, int out_space, int col_major_out) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    if (n_images < 1 || res_factor < 1) return fail(SF_E_INVALID, "n_images and res_factor must be >= 1");
    if (!bgr && !depth_raw) return fail(SF_E_INVALID, "no input image");
    if ((!bgr && (intensity || color_full)) || (!depth_raw && (depth || depth_mm))) return fail(SF_E_INVALID, "output requested without its input");
    if (col_major_out && out_space == SF_MEM_DEVICE) return fail(SF_E_INVALID, "column-major output is a host-side conversion");
    CU(cudaSetDevice(c->device));
    const size_t P = c->a.P0, n = (size_t)n_images, Pf = P * res_factor * res_factor;
    const bool need_in = in_space != SF_MEM_DEVICE, need_out = out_space != SF_MEM_DEVICE;
    // scratch layout: [bgr in][depth in][intensity][depth][depth_mm][colour]
    const size_t o_bgr = 0, o_raw = o_bgr + align256(need_in && bgr ? 3 * Pf * n : 0), o_i = o_raw + align256(need_in && depth_raw ? 2 * Pf * n : 0);
    const size_t o_d = o_i + align256(need_out && intensity ? 4 * P * n : 0), o_mm = o_d + align256(need_out && depth ? 4 * P * n : 0);
    const size_t o_col = o_mm + align256(need_out && depth_mm ? 2 * P * n : 0), total = o_col + align256(need_out && color_full ? 3 * P * n : 0);
    if (total) { const int rc = cvt_reserve(c, total); if (rc) return rc; }
    const uint8_t* s_bgr = bgr;
    const uint16_t* s_raw = depth_raw;
    if (need_in && bgr) { CU(cudaMemcpyAsync(c->d_cvt + o_bgr, bgr, 3 * Pf * n, cudaMemcpyHostToDevice, c->stream)); s_bgr = c->d_cvt + o_bgr; }
    if (need_in && depth_raw) { CU(cudaMemcpyAsync(c->d_cvt + o_raw, depth_raw, 2 * Pf * n, cudaMemcpyHostToDevice, c->stream)); s_raw = (const uint16_t*)(c->d_cvt + o_raw); }
    float* t_i = intensity ? (need_out ? (float*)(c->d_cvt + o_i) : intensity) : nullptr;
    float* t_d = depth ? (need_out ? (float*)(c->d_cvt + o_d) : depth) : nullptr;
    uint16_t* t_mm = depth_mm ? (need_out ? (uint16_t*)(c->d_cvt + o_mm) : depth_mm) : nullptr;
    uint8_t* t_col = color_full ? (need_out ? c->d_cvt + o_col : color_full) : nullptr;
    launch_convert_frames(s_bgr, s_raw, c->p.rows, c->p.cols, res_factor, n_images, t_i, t_d, P, t_mm, t_col, c->stream);
    CU(cudaGetLastError());
    if (need_out) {
        std::vector<float> tmp;
        const int rows = c->p.rows, cols = c->p.cols;
        auto fetch_f32 = [&](float* dst, const float* dev) -> int {
            if (!col_major_out) { CU(cudaMemcpyAsync(dst, dev, 4 * P * n, cudaMemcpyDeviceToHost, c->stream)); return SF_OK; }
            tmp.resize(P * n);
            CU(cudaMemcpyAsync(tmp.data(), dev, 4 * P * n, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            for (size_t k = 0; k < n; k++)
                for (int v = 0; v < rows; v++)
                    for (int u = 0; u < cols; u++) dst[k * P + (size_t)u * rows + v] = tmp[k * P + (size_t)v * cols + u];
            return SF_OK;
        };
        int rc;
        if (intensity && (rc = fetch_f32(intensity, t_i))) return rc;
        if (depth && (rc = fetch_f32(depth, t_d))) return rc;
        if (depth_mm) CU(cudaMemcpyAsync(depth_mm, t_mm, 2 * P * n, cudaMemcpyDeviceToHost, c->stream));
        if (color_full) CU(cudaMemcpyAsync(color_full, t_col, 3 * P * n, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return SF_OK;
}

int sf_upload_sequence_raw(sf_ctx* c, int n_frames, const uint8_t* bgr, const uint16_t* depth_raw, int res_factor, int in_space,
                           const float* twist_old_in) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    if (n_frames < 2 || n_frames - 1 > c->max_batch) return fail(SF_E_INVALID, "n_frames-1 must be in [1, max_batch]");
    if (!bgr || !depth_raw) return fail(SF_E_INVALID, "NULL image pointer");
    if (res_factor < 1) return fail(SF_E_INVALID, "res_factor must be >= 1");
    CU(cudaSetDevice(c->device));
    const size_t P = c->a.P0, n = (size_t)n_frames, Pf = P * res_factor * res_factor;
    const uint8_t* s_bgr = bgr;
    const uint16_t* s_raw = depth_raw;
    if (in_space != SF_MEM_DEVICE) {
        const size_t o_raw = align256(3 * Pf * n);
        const int rc = cvt_reserve(c, o_raw + align256(2 * Pf * n));
        if (rc) return rc;
        cudaStream_t st = c->stream;
        if (c->copy_streams) {  // the frames travel on the upload stream while this context (and the others) solve
            st = c->ul_stream;
            if (c->cvt_recorded) CU(cudaStreamWaitEvent(st, c->ev_cvt, 0));  // the last conversion has read the scratch
        }
        CU(cudaMemcpyAsync(c->d_cvt, bgr, 3 * Pf * n, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(c->d_cvt + o_raw, depth_raw, 2 * Pf * n, cudaMemcpyHostToDevice, st));
        if (c->copy_streams) { CU(cudaEventRecord(c->ev_ul, st)); CU(cudaStreamWaitEvent(c->stream, c->ev_ul, 0)); }
        s_bgr = c->d_cvt; s_raw = (const uint16_t*)(c->d_cvt + o_raw);
    }
    // frame k converts straight into the level-0 slot of its pyramids (prediction := previous raw frame)
    launch_convert_frames(s_bgr, s_raw, c->p.rows, c->p.cols, res_factor, n_frames, c->a.pyr_i, c->a.pyr_d, c->a.pyr_stride, nullptr, nullptr, c->stream);
    CU(cudaGetLastError());
    if (c->copy_streams && in_space != SF_MEM_DEVICE) { CU(cudaEventRecord(c->ev_cvt, c->stream)); c->cvt_recorded = true; }
    const int n_pairs = n_frames - 1;
    std::vector<int>&ci = c->h_ci, &pi = c->h_pi;
    if ((int)ci.size() != n_pairs || ci[0] != 1 || pi[0] != 0) {
        CU(cudaStreamSynchronize(c->stream));
        ci.resize(n_pairs); pi.resize(n_pairs);
        for (int k = 0; k < n_pairs; k++) { ci[k] = k + 1; pi[k] = k; }
    }
    CU(cudaMemcpyAsync(c->d_cur_idx, ci.data(), sizeof(int) * n_pairs, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_pred_idx, pi.data(), sizeof(int) * n_pairs, cudaMemcpyHostToDevice, c->stream));
    int rc;
    if ((rc = upload_twist(c, n_pairs, twist_old_in))) return rc;
    if (twist_old_in) CU(cudaStreamSynchronize(c->stream));
    c->n_pairs = n_pairs; c->n_frames = n_frames; c->uploaded = true; c->solved = false; c->is_sequence = true;
    return SF_OK;
}

int sf_buffer_push(sf_ctx* c, int index) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    if (!c->solved || c->n_pairs != 1) return fail(SF_E_STATE, "sf_buffer_push follows sf_run_solver (drop-in path, one pair)");
    CU(cudaSetDevice(c->device));
    const int b = ((index % 5) + 5) % 5;
    const size_t n = c->a.P0;
    // frame slot 0 = current frame of the drop-in path; level 0 is the head of its pyramid
    CU(cudaMemcpyAsync(c->a.ring_d + (size_t)b * n, c->a.pyr_d, sizeof(float) * n, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->a.ring_i + (size_t)b * n, c->a.pyr_i, sizeof(float) * n, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->a.ring_T + 16 * b, c->a.out[0].T, sizeof(float) * 16, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SF_OK;
}

int sf_compute_residuals_against_previous_image(sf_ctx* c, int index) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    if (!c->solved || c->n_pairs != 1) return fail(SF_E_STATE, "runSolver has not run (drop-in path, one pair)");
    if (index < 5) return fail(SF_E_INVALID, "needs im_count >= bufferLength = 5 (StaticFusion-datasets.cpp:175)");
    CU(cudaSetDevice(c->device));
    const LaunchCfg cfg{c->stream, 1, c->n_frames, nullptr};
    launch_history(c->a, c->dp, c->geom[0], 1, index, cfg);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return SF_OK;
}

int sf_get_outputs(sf_ctx* c, float T_odometry[16], float twist_old_out[6], float b_segm[SF_NUM_CLUSTERS], float* b_perpixel,
                   int32_t* labels, int col_major, int* irls_iterations, int* status) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    if (!c->solved) return fail(SF_E_STATE, "runSolver has not run");
    if (c->n_pairs != 1) return fail(SF_E_STATE, "sf_get_outputs reads the drop-in path's single pair; use sf_download after a batched solve");
    const int rows = c->p.rows, cols = c->p.cols;
    const size_t n = c->a.P0;
    std::vector<float> bp(b_perpixel ? n : 0);
    std::vector<uint8_t> lb(labels ? n : 0);
    const int rc = sf_download_range(c, 0, 1, T_odometry, twist_old_out, b_segm, b_perpixel ? bp.data() : nullptr, labels ? lb.data() : nullptr,
                                     SF_MEM_HOST, irls_iterations, status, nullptr);
    if (rc) return rc;
    for (int v = 0; v < rows; v++)
        for (int u = 0; u < cols; u++) {
            const size_t src = (size_t)v * cols + u;
            const size_t dst = col_major ? (size_t)u * rows + v : src;
            if (b_perpixel) b_perpixel[dst] = bp[src];
            if (labels) labels[dst] = (int32_t)lb[src];
        }
    return SF_OK;
}

// ---------------------------------------------------------------------------------------------
// measurement hooks
// ---------------------------------------------------------------------------------------------
int sf_profile_enable(sf_ctx* c, int on) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    c->prof_on = on != 0;
    return SF_OK;
}

int sf_profile_read(sf_ctx* c, float* ms, int* launches) {
    if (!c || !ms || !launches) return fail(SF_E_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < SF_PROF_CLASSES * SF_PROF_LEVELS; i++) { ms[i] = 0.f; launches[i] = 0; }
    for (const auto& r : c->prof) {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, r.e0, r.e1));
        ms[r.cls * SF_PROF_LEVELS + r.level] += t;
        launches[r.cls * SF_PROF_LEVELS + r.level] += 1;
    }
    return SF_OK;
}

int sf_profile_read_records(sf_ctx* c, int capacity, int* cls, int* level, float* ms, int* n_records) {
    if (!c || !cls || !level || !ms || !n_records || capacity < 0) return fail(SF_E_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    *n_records = (int)c->prof.size();
    for (int i = 0; i < (int)c->prof.size() && i < capacity; i++) {
        const auto& r = c->prof[i];
        CU(cudaEventElapsedTime(&ms[i], r.e0, r.e1));
        cls[i] = r.cls; level[i] = r.level;
    }
    return SF_OK;
}

int sf_get_step_stats(sf_ctx* c, int* n_valid, int* irls_iters) {
    if (!c || !n_valid || !irls_iters) return fail(SF_E_INVALID, "NULL argument");
    if (!c->solved) return fail(SF_E_STATE, "nothing has been solved");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    const size_t n = (size_t)c->n_pairs * c->a.trace_steps;
    std::vector<int> h(2 * n);
    CU(cudaMemcpy(h.data(), c->a.stepstat, sizeof(int) * 2 * n, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; i++) { n_valid[i] = h[2 * i]; irls_iters[i] = h[2 * i + 1]; }
    return SF_OK;
}

int sf_get_kmeans_iterations(sf_ctx* c, int* iterations) {
    if (!c || !iterations) return fail(SF_E_INVALID, "NULL argument");
    if (!c->solved) return fail(SF_E_STATE, "nothing has been solved");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    if (!c->p.enable_segmentation) { for (int k = 0; k < c->n_pairs; k++) iterations[k] = 0; return SF_OK; }
    CU(cudaMemcpy2D(iterations, sizeof(int), reinterpret_cast<const char*>(c->a.ctl) + offsetof(PairCtl, km_iters), sizeof(PairCtl), sizeof(int),
                    (size_t)c->n_pairs, cudaMemcpyDeviceToHost));
    return SF_OK;
}

int sf_result_rows_device(sf_ctx* c, const float** rows, int* n_rows, int* row_floats) {
    if (!c || !rows || !n_rows || !row_floats) return fail(SF_E_INVALID, "NULL argument");
    if (!c->solved) return fail(SF_E_STATE, "nothing has been solved");
    static_assert(sizeof(PairOut) == 48 * sizeof(float), "result row layout");
    *rows = reinterpret_cast<const float*>(c->a.out);
    *n_rows = c->n_pairs;
    *row_floats = (int)(sizeof(PairOut) / sizeof(float));
    return SF_OK;
}

// ---------------------------------------------------------------------------------------------
// introspection
// ---------------------------------------------------------------------------------------------
int sf_debug_set_stop_step(sf_ctx* c, int stop_step) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    c->stop_step = stop_step;
    return SF_OK;
}

int sf_debug_get_plane(sf_ctx* c, const char* name, int pair, int image_level, float* out) {
    if (!c || !name || !out) return fail(SF_E_INVALID, "NULL argument");
    if (pair < 0 || pair >= c->n_pairs || image_level < 0 || image_level >= c->levels) return fail(SF_E_INVALID, "pair / level out of range");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    const LevelGeom& g = c->geom[image_level];
    const Arena& a = c->a;
    int ci = 0, pi = 0;
    CU(cudaMemcpy(&ci, c->d_cur_idx + pair, sizeof(int), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&pi, c->d_pred_idx + pair, sizeof(int), cudaMemcpyDeviceToHost));
    const std::string n(name);
    const float* src = nullptr;
    if (n == "depth") src = a.pyr_d + (size_t)ci * a.pyr_stride + g.off;
    else if (n == "intensity") src = a.pyr_i + (size_t)ci * a.pyr_stride + g.off;
    else if (n == "depth_pred") src = a.pyr_d + (size_t)pi * a.pyr_stride + g.off;
    else if (n == "intensity_pred") src = a.pyr_i + (size_t)pi * a.pyr_stride + g.off;
    else if (n == "depth_warped") src = a.warp_d + (size_t)pair * a.P0;
    else if (n == "intensity_warped") src = a.warp_i + (size_t)pair * a.P0;
    else if (n == "depth_warped_ref" && image_level == 0) src = a.warp_d + (size_t)pair * a.P0;      // after the history stage
    else if (n == "intensity_warped_ref" && image_level == 0) src = a.warp_i + (size_t)pair * a.P0;
    else {
        static const char* names[NPLANES] = {"depth_inter", "xx_inter", "yy_inter", "dcu", "dcv", "dct", "ddu", "ddv", "ddt", "weights_c", "weights_d"};
        for (int k = 0; k < NPLANES; k++)
            if (n == names[k]) {
                if (!a.dbg) return fail(SF_E_STATE, "linearisation planes are only kept when the context was created with the trace flag");
                src = a.dbg + ((size_t)pair * NPLANES + k) * a.P0;
            }
    }
    if (src) {
        CU(cudaMemcpy(out, src, sizeof(float) * g.P, cudaMemcpyDeviceToHost));
        return SF_OK;
    }
    if (n == "valid") {
        const size_t nt = tiles_per_pair((size_t)g.P);
        std::vector<uint8_t> v(nt * TILE_BYTES);
        CU(cudaMemcpy(v.data(), a.tiles + (size_t)pair * tiles_per_pair(a.P0) * TILE_BYTES, v.size(), cudaMemcpyDeviceToHost));
        for (int i = 0; i < g.P; i++) out[i] = (v[tile_label_off(i)] != VLABEL_INVALID) ? 1.f : 0.f;
        return SF_OK;
    }
    return fail(SF_E_INVALID, "unknown plane name");
}

int sf_debug_get_labels(sf_ctx* c, int pair, int image_level, int32_t* out) {
    if (!c || !out) return fail(SF_E_INVALID, "NULL argument");
    if (pair < 0 || pair >= c->n_pairs || image_level < 0 || image_level >= c->levels) return fail(SF_E_INVALID, "pair / level out of range");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    const LevelGeom& g = c->geom[image_level];
    std::vector<uint8_t> v(g.P);
    CU(cudaMemcpy(v.data(), c->a.labels + (size_t)pair * c->a.pyr_stride + g.off, g.P, cudaMemcpyDeviceToHost));
    for (int i = 0; i < g.P; i++) out[i] = v[i];
    return SF_OK;
}

int sf_debug_get_kmeans(sf_ctx* c, int pair, float centres[3 * SF_NUM_CLUSTERS], uint8_t connectivity[SF_NUM_CLUSTERS * SF_NUM_CLUSTERS]) {
    if (!c) return fail(SF_E_INVALID, "ctx is NULL");
    if (pair < 0 || pair >= c->n_pairs) return fail(SF_E_INVALID, "pair out of range");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    PairCtl h;
    CU(cudaMemcpy(&h, c->a.ctl + pair, sizeof(PairCtl), cudaMemcpyDeviceToHost));
    if (centres) std::memcpy(centres, h.kmeans, sizeof(float) * 3 * NC);
    if (connectivity)
        for (int i = 0; i < NC; i++)
            for (int j = 0; j < NC; j++) connectivity[i * NC + j] = (h.conn[i] >> j) & 1u;
    return SF_OK;
}

int sf_debug_get_trace(sf_ctx* c, int pair, float* out, int n_floats) {
    if (!c || !out) return fail(SF_E_INVALID, "NULL argument");
    if (!c->a.trace) return fail(SF_E_STATE, "context was created without the trace flag");
    if (pair < 0 || pair >= c->n_pairs) return fail(SF_E_INVALID, "pair out of range");
    const int need = c->a.trace_steps * SF_TRACE_STEP;
    if (n_floats < need) return fail(SF_E_INVALID, "trace buffer too small");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(out, c->a.trace + (size_t)pair * need, sizeof(float) * need, cudaMemcpyDeviceToHost));
    return SF_OK;
}

}  // extern "C"
