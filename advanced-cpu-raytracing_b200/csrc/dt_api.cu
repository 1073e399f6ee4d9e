// C ABI of the render hot path (include/dorktracer.h): scene upload, wavefront driver, outputs.
#include "dt_flatten.h"
#include "dt_flatten_gpu.h"
#include "dt_kernels.cuh"
#include "../../include/dorktracer_debug.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;
int g_device = -1;
thread_local int tl_device = -1;          // dt_multi_create: the device a worker thread creates its scene on (overrides g_device)

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_err = std::string(#call) + ": " + cudaGetErrorString(e_); return DT_ERR_CUDA; } } while (0)

int ensure_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        g_err = std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") + "); this library has no CPU fallback";
        return DT_ERR_NO_DEVICE;
    }
    if (g_device < 0) g_device = 0;
    const int dev = tl_device >= 0 ? tl_device : g_device;
    if (dev >= n) { g_err = "CUDA device index out of range"; return DT_ERR_NO_DEVICE; }
    e = cudaSetDevice(dev);
    if (e != cudaSuccess) { g_err = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return DT_ERR_CUDA; }
    return DT_OK;
}

template <class T>
int upload(std::vector<void*>& allocs, const T* src, size_t n, const T** dst, size_t pad_elems = 0) {
    *dst = nullptr;
    size_t bytes = std::max<size_t>((n + pad_elems) * sizeof(T), 16);
    void* p = nullptr;
    CK(cudaMalloc(&p, bytes));
    allocs.push_back(p);
    CK(cudaMemset(p, 0, bytes));
    if (n) CK(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    *dst = (const T*)p;
    return DT_OK;
}

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    bool armed = false;
    int init() { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); return DT_OK; }
    void start(cudaStream_t s) { cudaEventRecord(a, s); }
    void stop(cudaStream_t s) { cudaEventRecord(b, s); armed = true; }
    float take() { if (!armed) return 0.f; float ms = 0.f; cudaEventElapsedTime(&ms, a, b); armed = false; return ms; }
    void destroy() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); a = b = nullptr; }
};

}  // namespace

// for the other translation units of the library (dt_build.cu)
void dt_internal_set_error(const std::string& msg) { g_err = msg; }
int dt_internal_ensure_device() { return ensure_device(); }


#define DT_MAX_PIPES 8

struct DtPipe {
    DtRayQueue q[2];
    float4* miss[2] = {nullptr, nullptr};
    DtShadowQueue sq[3];          // wave k fills sq[k % 3]; shadow(k) on stream B overlaps closest(k+1) / shade(k+1) on stream A
    int capacity = 0, shadow_capacity = 0;
    bool has_miss = false, has_defer = false;
    std::vector<void*> allocs;
    cudaStream_t A = nullptr, B = nullptr;
    cudaEvent_t ev_shade[3] = {nullptr, nullptr, nullptr}, ev_shadow[3] = {nullptr, nullptr, nullptr}, ev_done = nullptr;
    int* counters = nullptr;      // this pipe's block of dt_scene::counters
    int* sort_perm = nullptr;     // material-sorted order of the current wave (sort stage)
    int* sort_hist = nullptr;     // DT_SORT_BINS bin counts + DT_SORT_BINS running cursors
    DtSpatialSort ss = {};        // counter tables of the hit-cell sort (k_ssort_*)
    DtPipe() { memset(q, 0, sizeof q); memset(sq, 0, sizeof sq); }
    void free_queues() { for (void* p : allocs) cudaFree(p); allocs.clear(); capacity = shadow_capacity = 0; }
};

struct dt_scene {
    int device = 0;
    int num_sms = 148;
    size_t n_blas_nodes = 0, n_prims = 0, n_faces = 0;     // sizes of the acceleration arrays (dt_scene_accel_checksum)
    cudaStream_t stream = nullptr;
    DtSceneDev dev;
    std::vector<void*> allocs;
    uint64_t n_triangles = 0;
    int lights_shadowed = 0;          // shadow rays per shaded hit
    int fanout_hint = 0;
    bool has_env = false;
    bool sphere_normal_maps = false;  // a sphere carries a normal map: k_shade<true> / k_tail<true> (dt_stale_normal)

    // render-time buffers: up to DT_MAX_PIPES independent wavefront pipelines (queues, counters, two streams each).
    // The frame's tiles are dealt round-robin to the pipelines; their waves run concurrently so that the tail of one
    // pipeline's persistent traversal kernel is filled by the bulk of another's.  Pipe 0 also serves the
    // host-synchronised loop (path tracing with Russian roulette / multi-batch frames).
    DtPipe pipes[DT_MAX_PIPES];
    cudaEvent_t ev_fork = nullptr;
    std::vector<cudaEvent_t> ev_pool;    // timing events of the sync-free loop
    int use_graph = 0;            // DT_GRAPH=1: replay the sync-free frame as a CUDA graph instead of enqueueing every launch (see the wave loop)
    bool capturing = false;
    cudaGraphExec_t frame_graph = nullptr;
    std::string frame_key;        // everything baked into frame_graph
    std::vector<cudaEvent_t> g_tg, g_tc, g_th, g_ts, g_tsort;
    uint32_t g_launches = 0, g_closest = 0;
    int grid_shade = 0;
    int sync_waves = 0;           // DT_SYNC_WAVES=1 forces the host-synchronised wave loop (A/B)
    int n_pipes_env = 0;          // DT_PIPES=n forces the pipeline count (0 = auto)
    int debug_timing = 0;         // DT_DEBUG_TIMING=1 prints host-side enqueue times to stderr
    int* counters = nullptr;      // DT_MAX_PIPES x DT_CNT_COUNT device ints (pipe p owns block p)
    int* h_counters = nullptr;    // pinned mirror (+ scratch)
    float4* accum = nullptr; size_t accum_pix = 0;
    float* hdr = nullptr; uint8_t* ldr = nullptr; size_t out_pix = 0;
    float* peer_hdr = nullptr; uint8_t* peer_ldr = nullptr;     // another rank's frame buffers (dt_frame_import) / device 0's (dt_multi)
    int peer_w = 0, peer_h = 0;
    bool peer_is_ipc = false;     // the peer pointers are CUDA IPC mappings (closed by dt_frame_release)
    bool frame_exported = false;  // other processes hold IPC mappings of hdr / ldr: they must not be reallocated
    double* tm_logsum = nullptr; unsigned int* tm_hist = nullptr; unsigned long long* tm_rank = nullptr; uint32_t* tm_prefix = nullptr;
    Timer t_total, t_gen, t_closest, t_shadow, t_shade, t_sort, t_resolve, t_tm;
    int grid_trav[4][2] = {};
    int trav_mode = 2, refill_threshold = 12;
    int shadow_spare = 1;         // persistent blocks per SM the overlapped shadow pass leaves free, so that the sort / shade / closest launches of
                                  // the critical path find SM slots while it runs (r1e_ab_shadow_order.log: 3.90 ms against 4.03 ms with 0)
    int shadow_order = 1;         // 1: shadow(k) released together with closest(k+1) (see the wave loop); 0: right after shade(k)
    int sort_mode = 1;            // sort-by-material stage: 0 off, 1 auto (scenes with >= 3 materials), 2 always (DT_SORT)
    int sort_spatial = 0;         // hit-cell sort (k_ssort_*) instead of the material sort: 0 off (default: measured slower, profiles/r2_ab_spatial_sort.log),
                                  // -1 path-traced frames only, 1..6 every frame, with at most this many cell bits per axis (DT_SORT_SPATIAL)
    int sort_refine = 1;          // second level of the hit-cell sort (DT_SORT_REFINE)

    // device-resident wave loop (path tracing with Russian roulette, multi-batch frames): one CUDA graph per frame shape
    int dev_loop = 1;             // DT_DEVLOOP=0: host-synchronised loop instead (A/B)
    int tail_threshold = -2;      // rays alive at which the loop hands over to k_tail; -2 = auto (64 per k_tail block), -1 = never (DT_TAIL_THRESHOLD)
    cudaGraph_t loop_graph = nullptr;
    cudaGraphExec_t loop_exec = nullptr;
    std::string loop_key;
    uint32_t loop_launches_per_iter = 0;
    DtTailMem tail; int tail_grid = 0; bool tail_has_miss = false, tail_has_defer = false;
    std::vector<void*> tail_allocs;

    void free_queues() { for (DtPipe& p : pipes) p.free_queues(); }
    void free_loop() {
        if (loop_exec) cudaGraphExecDestroy(loop_exec);
        if (loop_graph) cudaGraphDestroy(loop_graph);
        loop_exec = nullptr; loop_graph = nullptr; loop_key.clear();
    }
    void free_tail() { for (void* p : tail_allocs) cudaFree(p); tail_allocs.clear(); tail_grid = 0; }
};

namespace {

template <bool ANY>
void launch_traverse(dt_scene* s, const DtRayQueue& q, const DtShadowQueue& sq, const int* n_ptr, int n_fixed, int n_cap, int* fetch, float4* accum, cudaStream_t st = nullptr, int spare_blocks_per_sm = 0) {
    const int grid = std::max(s->num_sms, s->grid_trav[s->trav_mode][ANY ? 1 : 0] - spare_blocks_per_sm * s->num_sms);
    if (!st) st = s->stream;
    switch (s->trav_mode) {
        case 0: k_traverse<ANY, false><<<grid, 128, 0, st>>>(s->dev, q, sq, n_ptr, n_fixed, n_cap, fetch, accum); break;
        case 1: k_traverse<ANY, true><<<grid, 128, 0, st>>>(s->dev, q, sq, n_ptr, n_fixed, n_cap, fetch, accum); break;
        case 2: k_traverse_dyn<ANY, false><<<grid, 128, 0, st>>>(s->dev, q, sq, n_ptr, n_fixed, n_cap, fetch, accum, s->refill_threshold); break;
        default: k_traverse_dyn<ANY, true><<<grid, 128, 0, st>>>(s->dev, q, sq, n_ptr, n_fixed, n_cap, fetch, accum, s->refill_threshold); break;
    }
}

template <class T>
int qalloc(DtPipe& pp, T** p, size_t n) {
    void* v = nullptr;
    CK(cudaMalloc(&v, std::max<size_t>(n * sizeof(T), 16)));
    pp.allocs.push_back(v);
    *p = (T*)v;
    return DT_OK;
}

int ensure_queues(dt_scene* s, DtPipe& pp, int capacity, int shadow_capacity, bool need_defer) {
    if (pp.capacity >= capacity && pp.shadow_capacity >= shadow_capacity && pp.has_miss == s->has_env && (pp.has_defer || !need_defer)) return DT_OK;
    pp.free_queues();
    int rc;
    for (int k = 0; k < 2; k++) {
        DtRayQueue& q = pp.q[k];
        if ((rc = qalloc(pp, &q.o_time, capacity))) return rc;
        if ((rc = qalloc(pp, &q.d_tmax, capacity))) return rc;
        if ((rc = qalloc(pp, &q.hit0, capacity))) return rc;
        if ((rc = qalloc(pp, &q.hit_face, capacity))) return rc;
        if ((rc = qalloc(pp, &q.pixel, capacity))) return rc;
        if ((rc = qalloc(pp, &q.weight_n, capacity))) return rc;
        if ((rc = qalloc(pp, &q.thr_beer, capacity))) return rc;
        if ((rc = qalloc(pp, &q.misc, capacity))) return rc;
        if ((rc = qalloc(pp, &q.sort_key, capacity))) return rc;
        pp.miss[k] = nullptr;
        if (s->has_env) { if ((rc = qalloc(pp, &pp.miss[k], capacity))) return rc; }
    }
    for (int k = 0; k < 3; k++) {
        DtShadowQueue& sq = pp.sq[k];
        if ((rc = qalloc(pp, &sq.o_time, shadow_capacity))) return rc;
        if ((rc = qalloc(pp, &sq.d_tmax, shadow_capacity))) return rc;
        if ((rc = qalloc(pp, &sq.contrib_pix, shadow_capacity))) return rc;
        sq.defer = nullptr;
        if (need_defer && k == 0) { if ((rc = qalloc(pp, &sq.defer, shadow_capacity))) return rc; }
    }
    if ((rc = qalloc(pp, &pp.sort_perm, capacity))) return rc;
    if ((rc = qalloc(pp, &pp.sort_hist, 2 * DT_SORT_BINS))) return rc;
    pp.ss = DtSpatialSort();
    if (s->sort_spatial != 0) {                      // counter tables of the opt-in hit-cell sort
        if ((rc = qalloc(pp, &pp.ss.hist, DT_SSORT_MAX_BINS)) || (rc = qalloc(pp, &pp.ss.cursor, DT_SSORT_MAX_BINS)) || (rc = qalloc(pp, &pp.ss.blocksum, DT_SSORT_MAX_BINS / DT_SSORT_SCAN_BLOCK))) return rc;
        CK(cudaMemset(pp.ss.hist, 0, DT_SSORT_MAX_BINS * sizeof(int)));
    }
    pp.capacity = capacity; pp.shadow_capacity = shadow_capacity; pp.has_miss = s->has_env; pp.has_defer = need_defer;
    return DT_OK;
}

int ensure_outputs(dt_scene* s, size_t n_pix) {
    if (s->accum_pix < n_pix) {
        if (s->frame_exported) { g_err = "the frame buffers are exported to other processes (dt_frame_export) and cannot grow; export a frame of the largest resolution first"; return DT_ERR_INVALID; }
        if (s->accum) cudaFree(s->accum);
        if (s->hdr) cudaFree(s->hdr);
        if (s->ldr) cudaFree(s->ldr);
        s->accum = nullptr; s->hdr = nullptr; s->ldr = nullptr; s->accum_pix = 0;
        CK(cudaMalloc(&s->accum, n_pix * sizeof(float4)));
        CK(cudaMalloc(&s->hdr, n_pix * 3 * sizeof(float)));
        CK(cudaMalloc(&s->ldr, n_pix * 3));
        s->accum_pix = n_pix;
    }
    return DT_OK;
}

DtCamDev make_cam(const dt_camera_desc* c, int flags) {
    DtCamDev d;
    memset(&d, 0, sizeof d);                 // the struct is part of the graph cache keys: no indeterminate padding
    memcpy(d.position, c->position, 12); memcpy(d.gaze, c->gaze, 12); memcpy(d.up, c->up, 12); memcpy(d.right, c->right, 12); memcpy(d.q, c->q, 12);
    d.left = c->left; d.right_ = c->right_; d.bottom = c->bottom; d.top = c->top;
    d.width = c->width; d.height = c->height; d.spp = c->samples_per_pixel < 1 ? 1 : c->samples_per_pixel;
    d.focus_distance = c->focus_distance; d.aperture_size = c->aperture_size;
    d.path_tracing = c->path_tracing; d.importance_sampling = c->importance_sampling; d.nee = c->next_event_estimation; d.russian_roulette = c->russian_roulette;
    d.jitter_aa = (flags & DT_FLAG_JITTER_AA) ? 1 : 0;
    d.keep_weightless = (flags & DT_FLAG_KEEP_WEIGHTLESS_PATHS) ? 1 : 0;
    d.smooth_shading = (flags & DT_FLAG_SMOOTH_SHADING) ? 1 : 0;
    d.row_limit = (flags & DT_FLAG_REF_ROW_BANDS) ? (c->height / 8) * 8 : c->height;      // main.cpp:15,38-39: 8 bands of H / 8 rows
    return d;
}

int check_cam(const dt_camera_desc* c) {
    if (!c) { g_err = "null camera"; return DT_ERR_INVALID; }
    if (c->width <= 0 || c->height <= 0 || (long long)c->width * c->height > 0x0FFFFFFF) { g_err = "bad image resolution"; return DT_ERR_INVALID; }
    return DT_OK;
}

// Tonemap a device radiance buffer into a device LDR buffer.
int tonemap_device(dt_scene* s, const float* hdr, int W, int H, float key, float burn, float sat, float gamma, uint8_t* ldr, uint32_t* launches) {
    const int n_pix = W * H;
    const size_t n_vals = (size_t)n_pix * 3;
    cudaStream_t st = s->stream;
    CK(cudaMemsetAsync(s->tm_logsum, 0, sizeof(double), st));
    k_tm_logsum<<<s->num_sms * 4, 256, 0, st>>>(hdr, n_pix, s->tm_logsum);
    (*launches)++;
    CK(cudaMemsetAsync(s->tm_prefix, 0, sizeof(uint32_t), st));
    if (burn > 0.01) {
        // index of the white point in the sorted list of all channel values (tonemapper.h:100-104): evaluated in
        // FLOAT, (int)(thresholdPerct * lastIdx), so it is itself rounded above 2^24.
        float thresholdPerct = (100.0f - burn) / 100;
        int lastIdx = (int)n_vals - 1;
        int bi = (int)(thresholdPerct * lastIdx);
        if (bi > lastIdx) bi = lastIdx;
        if (bi < 0) bi = 0;
        unsigned long long rank = (unsigned long long)bi;
        CK(cudaMemcpyAsync(s->tm_rank, &rank, sizeof rank, cudaMemcpyHostToDevice, st));       // pageable source: staged before the call returns
        CK(cudaMemsetAsync(s->tm_hist, 0, 256 * sizeof(unsigned int), st));
        uint32_t mask = 0;
        for (int pass = 0; pass < 4; pass++) {
            const int shift = 24 - 8 * pass;
            k_tm_hist<<<s->num_sms * 8, 256, 0, st>>>(hdr, n_vals, s->tm_prefix, mask, shift, s->tm_hist);
            k_tm_pick<<<1, 1, 0, st>>>(s->tm_hist, s->tm_rank, s->tm_prefix, shift);
            *launches += 2;
            mask |= 0xFFu << shift;
        }
    }
    k_tm_map<<<(n_pix + 255) / 256, 256, 0, st>>>(hdr, n_pix, s->tm_logsum, s->tm_prefix, key, burn, sat, gamma, ldr);
    (*launches)++;
    CK(cudaGetLastError());
    return DT_OK;
}

struct RenderOut { float* hdr_dev; };

// Sort stage (k_sort_*): fills pp.sort_perm with the material-sorted order of wave queue `q`; returns launches.
int launch_sort(dt_scene* s, DtPipe& pp, const DtRayQueue& q, const int* n_ptr, int n_fixed, cudaStream_t st, bool zero_hist = true, int spatial_bits = 0) {
    if (spatial_bits > 0) {                          // by hit cell, then material (see k_ssort_hist)
        DtSpatialSort ss = pp.ss;
        ss.max_axis_bits = std::min(spatial_bits, DT_SSORT_MAX_AXIS_BITS); ss.refine = s->sort_refine;
        const int grid = s->num_sms * 8, scan_grid = DT_SSORT_MAX_BINS / DT_SSORT_SCAN_BLOCK;
        k_ssort_hist<<<grid, 256, 0, st>>>(s->dev, q, n_ptr, n_fixed, ss);
        k_ssort_scan_a<<<scan_grid, 256, 0, st>>>(n_ptr, n_fixed, ss);
        k_ssort_scan_b<<<scan_grid, 256, 0, st>>>(n_ptr, n_fixed, ss);
        k_ssort_scatter<<<grid, 256, 0, st>>>(q, n_ptr, n_fixed, ss, pp.sort_perm);
        if (ss.refine) k_ssort_refine<<<grid, 256, 0, st>>>(q, n_ptr, n_fixed, pp.sort_perm);
        return ss.refine ? 5 : 4;
    }
    if (zero_hist) cudaMemsetAsync(pp.sort_hist, 0, 2 * DT_SORT_BINS * sizeof(int), st);           // bin counts + running cursors (the device loop zeroes them in k_loop_begin)
    const int grid = s->num_sms * 4;
    k_sort_hist<<<grid, 256, 0, st>>>(s->dev, q, n_ptr, n_fixed, pp.sort_hist);
    k_sort_scatter<<<grid, 256, 0, st>>>(q, n_ptr, n_fixed, pp.sort_hist, pp.sort_hist + DT_SORT_BINS, pp.sort_perm);
    return 2;
}

// valid primary rays of this rank: every in-image pixel of the owned tiles
long long count_valid_pixels(const dt_render_params& P, int W, int H_image, int row_limit) {
    const int tiles_x = (W + 7) / 8, tiles_y = (H_image + 3) / 4;
    const int H = std::min(H_image, row_limit);                 // rows that get camera rays
    if (P.tile_world == 1 && W % 8 == 0 && H % 4 == 0) return (long long)W * H;
    long long valid = 0;
    const long long mine = dt_rank_tile_count(tiles_x, tiles_y, P.tile_rank, P.tile_world);
    for (long long j = 0; j < mine; j++) {
        int tx, ty;
        if (!dt_rank_tile(j, P.tile_rank, P.tile_world, tiles_x, tiles_y, tx, ty)) continue;
        int w = std::min(8, W - tx * 8), h = std::min(4, H - ty * 4);
        if (w > 0 && h > 0) valid += (long long)w * h;
    }
    return valid;
}


template <class T>
int talloc(std::vector<void*>& allocs, T** p, size_t n) {
    void* v = nullptr;
    CK(cudaMalloc(&v, std::max<size_t>(n * sizeof(T), 16)));
    allocs.push_back(v);
    *p = (T*)v;
    return DT_OK;
}

// Block-private queues of k_tail: `grid` blocks x `cap` rays (+ 3 x cap x shadows_per_hit shadow rays).
int ensure_tail(dt_scene* s, int shadows_per_hit, bool need_defer) {
    if (s->tail_grid > 0 && s->tail_has_miss == s->has_env && (s->tail_has_defer || !need_defer) && s->tail.shadow_capacity >= s->tail.capacity * shadows_per_hit) return DT_OK;
    s->free_tail();
    s->free_loop();                      // the graph holds the old pointers
    int bps = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_tail<false>, DT_TAIL_THREADS, 0));
    const int grid = s->num_sms * std::max(1, bps);
    DtTailMem& M = s->tail;
    memset(&M, 0, sizeof M);
    M.capacity = 1024; M.shadow_capacity = M.capacity * shadows_per_hit;
    const size_t nq = (size_t)grid * M.capacity, ns = (size_t)grid * 3 * M.shadow_capacity;      // three rotating shadow buffers per block
    int rc;
#define A(pp_, n_) talloc(s->tail_allocs, pp_, n_)
    for (int k = 0; k < 2; k++) {
        DtRayQueue& q = M.q[k];
        if ((rc = A(&q.o_time, nq)) || (rc = A(&q.d_tmax, nq)) || (rc = A(&q.hit0, nq)) || (rc = A(&q.hit_face, nq)) || (rc = A(&q.pixel, nq)) ||
            (rc = A(&q.weight_n, nq)) || (rc = A(&q.thr_beer, nq)) || (rc = A(&q.misc, nq))) return rc;
        q.sort_key = nullptr;
        if (s->has_env) { if ((rc = A(&M.miss[k], nq))) return rc; }
    }
    if ((rc = A(&M.sq.o_time, ns)) || (rc = A(&M.sq.d_tmax, ns)) || (rc = A(&M.sq.contrib_pix, ns))) return rc;
    if (need_defer) { if ((rc = A(&M.sq.defer, ns))) return rc; }
#undef A
    s->tail_grid = grid; s->tail_has_miss = s->has_env; s->tail_has_defer = need_defer;
    return DT_OK;
}

// Device-resident wave loop (see k_loop_begin in dt_kernels.cuh): the whole frame is one graph launch and one host sync.
// *overflow: a queue overflowed on the device (the frame is incomplete; the caller retries with smaller waves).
int render_devloop(dt_scene* s, const DtCamDev& dc, const DtWaveParams& wp, long long total, int wave_max, int capacity, int shadow_capacity,
                   int shadows_per_hit, bool defer_mode, bool do_sort, int sort_spatial, dt_stats& S, bool* overflow) {
    DtPipe& pp = s->pipes[0];
    int rc;
    if ((rc = ensure_queues(s, pp, capacity, shadow_capacity, defer_mode))) return rc;
    const bool use_tail = s->tail_threshold != -1;
    if (use_tail && (rc = ensure_tail(s, shadows_per_hit, defer_mode))) return rc;
    const int tail_threshold = !use_tail ? -1 : (s->tail_threshold >= 0 ? std::min(s->tail_threshold, s->tail_grid * (s->tail.capacity / 4)) : s->tail_grid * 64);
    cudaStream_t st = s->stream;
    int* c = pp.counters;
    DtShadowQueue& sq = pp.sq[0];

    std::string key;
    key.append((const char*)&dc, sizeof dc); key.append((const char*)&wp, sizeof wp);
    const long long misc[12] = {total, wave_max, pp.capacity, pp.shadow_capacity, defer_mode, do_sort + 2 * sort_spatial + 64 * s->sort_refine, tail_threshold, s->trav_mode, s->refill_threshold, s->dev.n_shapes, s->tail_grid, s->grid_shade};
    key.append((const char*)misc, sizeof misc);
    const void* ptrs[6] = {s->accum, c, pp.q[0].o_time, sq.o_time, pp.sort_perm, s->tail.q[0].o_time};
    key.append((const char*)ptrs, sizeof ptrs);

    if (!s->loop_exec || key != s->loop_key) {
        s->free_loop();
        cudaGraph_t g = nullptr;
        CK(cudaGraphCreate(&g, 0));
        s->loop_graph = g;
        cudaGraphConditionalHandle handle;
        CK(cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault));       // every launch starts with "run the body"
        cudaGraphNodeParams np = {};
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = handle; np.conditional.type = cudaGraphCondTypeWhile; np.conditional.size = 1;
        cudaGraphNode_t wnode = nullptr;
        CK(cudaGraphAddNode(&wnode, g, nullptr, 0, &np));
        cudaGraph_t body = np.conditional.phGraph_out[0];
        uint32_t launches = 0;
        CK(cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        for (int cur = 0; cur < 2; cur++) {
            k_loop_begin<<<1, 2 * DT_SORT_BINS, 0, st>>>(c, wave_max, total, do_sort ? pp.sort_hist : nullptr);
            k_generate_dev<<<s->num_sms * 8, 256, 0, st>>>(dc, wp, pp.q[cur], c, s->accum);
            launch_traverse<false>(s, pp.q[cur], sq, c + DT_CNT_CUR, 0, pp.capacity, c + DT_CNT_FETCH_A, s->accum, st);
            launches += 3;
            if (defer_mode) {
                k_filter_deferred_dev<<<s->num_sms * 4, 256, 0, st>>>(s->dev, sq, c + DT_CNT_PREV_SHADOW, pp.shadow_capacity, pp.q[cur]);
                launch_traverse<true>(s, pp.q[cur], sq, c + DT_CNT_PREV_SHADOW, 0, pp.shadow_capacity, c + DT_CNT_FETCH_B, s->accum, st);
                launches += 2;
            }
            if (do_sort) launches += launch_sort(s, pp, pp.q[cur], c + DT_CNT_CUR, 0, st, false, sort_spatial);
            {
                DtShadeCounters sc = {c + DT_CNT_NEXT, c + DT_CNT_SHADOW, c + DT_CNT_OVERFLOW, reinterpret_cast<unsigned long long*>(c + DT_CNT_SHADOW_DEAD), reinterpret_cast<unsigned long long*>(c + DT_CNT_CLOSEST_DEAD)};
                (s->sphere_normal_maps ? k_shade<true> : k_shade<false>)<<<s->grid_shade, 128, 0, st>>>(s->dev, dc, pp.q[cur], pp.miss[cur], c + DT_CNT_CUR, 0, do_sort ? pp.sort_perm : nullptr, pp.q[1 - cur], pp.miss[1 - cur], pp.capacity,
                                                        sq, pp.shadow_capacity, sc, s->accum);
                launches++;
            }
            if (!defer_mode) { launch_traverse<true>(s, pp.q[cur], sq, c + DT_CNT_SHADOW, 0, pp.shadow_capacity, c + DT_CNT_FETCH_B, s->accum, st); launches++; }
            k_loop_end<<<1, 1, 0, st>>>(c, pp.capacity, pp.shadow_capacity, defer_mode ? 1 : 0, total, tail_threshold, cur, handle);
            launches++;
        }
        cudaGraph_t captured = nullptr;
        cudaError_t ce = cudaStreamEndCapture(st, &captured);
        if (ce != cudaSuccess) { g_err = std::string("capture of the wave-loop body failed: ") + cudaGetErrorString(ce); s->free_loop(); return DT_ERR_CUDA; }
        if (use_tail) {
            CK(cudaStreamBeginCaptureToGraph(st, g, &wnode, nullptr, 1, cudaStreamCaptureModeThreadLocal));
            (s->sphere_normal_maps ? k_tail<true> : k_tail<false>)<<<s->tail_grid, DT_TAIL_THREADS, 0, st>>>(s->dev, dc, pp.q[0], pp.miss[0], sq, c, s->tail, defer_mode ? 1 : 0, s->accum);
            ce = cudaStreamEndCapture(st, &captured);
            if (ce != cudaSuccess) { g_err = std::string("capture of the tail kernel failed: ") + cudaGetErrorString(ce); s->free_loop(); return DT_ERR_CUDA; }
        }
        ce = cudaGraphInstantiate(&s->loop_exec, g, 0);
        if (ce != cudaSuccess) { s->loop_exec = nullptr; g_err = std::string("cudaGraphInstantiate (wave loop): ") + cudaGetErrorString(ce); s->free_loop(); return DT_ERR_CUDA; }
        s->loop_key = key;
        s->loop_launches_per_iter = launches;
    }
    CK(cudaGraphLaunch(s->loop_exec, st));
    CK(cudaMemcpyAsync(s->h_counters, c, DT_CNT_COUNT * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    const int* hc = s->h_counters;
    *overflow = hc[DT_CNT_OVERFLOW] != 0;
    unsigned long long c8 = 0, s8 = 0;
    unsigned long long dead8 = 0;
    memcpy(&c8, hc + DT_CNT_TOT_CLOSEST, 8); memcpy(&s8, hc + DT_CNT_TOT_SHADOW, 8); memcpy(&dead8, hc + DT_CNT_SHADOW_DEAD, 8);
    unsigned long long cdead8 = 0;
    memcpy(&cdead8, hc + DT_CNT_CLOSEST_DEAD, 8);
    S.rays_closest = c8 - std::min(c8, cdead8); S.rays_shadow = s8 - std::min(s8, dead8);          // queue entries minus the ones that are not traced
    S.waves = (uint32_t)(hc[DT_CNT_WAVES] + hc[DT_CNT_TAIL_WAVES]);
    S.kernel_launches = (uint32_t)hc[DT_CNT_ITERS] * s->loop_launches_per_iter + (use_tail ? 1u : 0u);
    S.launches_traverse_closest = (uint32_t)hc[DT_CNT_ITERS] * 2u;
    if (s->debug_timing) fprintf(stderr, "[dt] device loop: %d body iterations, %d waves + %d tail waves (%d rays handed to k_tail)\n", hc[DT_CNT_ITERS], hc[DT_CNT_WAVES], hc[DT_CNT_TAIL_WAVES], hc[DT_CNT_TAIL_RAYS]);
    return DT_OK;
}

int render_core(dt_scene* s, const dt_camera_desc* cam, const dt_render_params* params, dt_stats* stats, bool primary_only) {
    int rc = check_cam(cam);
    if (rc) return rc;
    CK(cudaSetDevice(s->device));
    dt_render_params P; memset(&P, 0, sizeof P); P.seed = 1234; P.tile_world = 1;
    if (params) P = *params;
    if (P.tile_world < 1 || P.tile_rank < 0 || P.tile_rank >= P.tile_world) { g_err = "bad tile_rank/tile_world"; return DT_ERR_INVALID; }
    DtCamDev dc = make_cam(cam, P.flags);
    if (primary_only) { dc.spp = 1; }
    const int W = cam->width, H = cam->height;
    const size_t n_pix = (size_t)W * H;
    if ((rc = ensure_outputs(s, n_pix))) return rc;

    DtWaveParams wp;
    wp.seed_lo = (uint32_t)P.seed; wp.seed_hi = (uint32_t)(P.seed >> 32);
    wp.tile_rank = P.tile_rank; wp.tile_world = P.tile_world;
    wp.tiles_x = (W + 7) / 8; wp.tiles_y = (H + 3) / 4;
    const long long my_tiles = dt_rank_tile_count(wp.tiles_x, wp.tiles_y, P.tile_rank, P.tile_world);
    wp.per_sample = my_tiles * 32;
    const long long total = wp.per_sample * dc.spp;

    const bool pt = dc.path_tracing != 0 && !primary_only;
    const bool defer_mode = pt && dc.nee && s->dev.n_mesh_lights > 0;
    const int fan = s->fanout_hint + (pt ? 1 : 0);
    const int shadows_per_hit = std::max(1, s->lights_shadowed);
    int wave_max = P.max_wave_rays;
    if (wave_max <= 0 && total <= (1ll << 23)) wave_max = 1 << 23;          // single-wave frames: nothing to size (cudaMemGetInfo costs ~1 ms, a quarter of a config-2 frame)
    if (wave_max <= 0) {
        // Default wave size: as large as HBM allows, up to 32 Mi rays.  Every wave ends in the drain of five persistent / grid-stride
        // launches, so frames of many waves (path tracing) run ~3 % faster with 32 Mi-ray waves than with 8 Mi (config 5:
        // profiles/r2_ab_wave_size.log); at ~2.2 KB of queue space per wave ray (4x fan-out, two shadow rays per hit) that is
        // 75 GB of the B200's 180 GB.  The wave shrinks by halves until the queues fit into 45 % of the free device memory.
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = 0;
        for (const DtPipe& pp : s->pipes) free_b += (size_t)pp.capacity * 256 + (size_t)pp.shadow_capacity * 152;      // what the current queues would give back (roughly)
        const double per_ray = (fan <= 1 ? 1.0 : 4.0) * (2 * 108 + (s->has_env ? 32 : 0) + 4 + shadows_per_hit * (3 * 48 + 8.0));
        wave_max = 1 << 25;
        const long long mult = fan <= 1 ? 1 : 4;
        while (wave_max > (1 << 20) && ((double)wave_max * per_ray > 0.45 * (double)free_b || mult * wave_max * shadows_per_hit > (1ll << 28))) wave_max >>= 1;   // 2^28: queue_caps' limit
    }
    wave_max = (int)std::min<long long>(wave_max, std::max<long long>(total, 32));
    wave_max = (wave_max + 31) & ~31;
    const bool sort_allowed = !primary_only && !(P.flags & DT_FLAG_NO_SORT) && s->sort_mode != 0;
    // hit-cell sort (k_ssort_*): path-traced frames by default (their waves are incoherent), any frame with DT_SORT_SPATIAL=1..6
    const int sort_spatial = (!sort_allowed || (P.flags & DT_FLAG_SORT_MATERIAL_ONLY)) ? 0 : (s->sort_spatial < 0 ? (pt ? DT_SSORT_MAX_AXIS_BITS : 0) : std::min(s->sort_spatial, DT_SSORT_MAX_AXIS_BITS));
    const bool do_sort = sort_allowed && (sort_spatial > 0 || s->sort_mode == 2 || (P.flags & DT_FLAG_FORCE_SORT) || s->dev.n_materials >= 3);
    const bool host_loop = s->sync_waves || (P.flags & (DT_FLAG_HOST_WAVE_LOOP | DT_FLAG_SERIAL_WAVES));
    const bool frame_graph = s->use_graph || (P.flags & DT_FLAG_FRAME_GRAPH);
    uint32_t retries = 0;

    cudaStream_t st = s->stream;
    dt_stats S; memset(&S, 0, sizeof S);
    const auto t_host0 = std::chrono::steady_clock::now();
    s->t_total.start(st);

    // Queue sizes for waves of `wave` rays.  A closest-hit wave can hold up to 4x its parents when the ray tree fans out (dielectric:
    // two children per hit; path tracing: GI child + specular child), and every one of those hits may be lit, so the shadow queue
    // follows the CLOSEST queue's capacity, not the wave size.  Every overflow retry shrinks the wave 4x and grows the ratio 4x.
    // DT_FLAG_TEST_TIGHT_QUEUES sizes the first attempt for a fan-out of 1 (test hook for the overflow -> retry path).
    auto queue_caps = [&](long long wave, long long& cap, long long& shcap) {
        const bool tight = (P.flags & DT_FLAG_TEST_TIGHT_QUEUES) && retries == 0;
        const long long mult = (fan <= 1 || tight) ? (retries == 0 ? 1ll : (1ll << (2 * retries))) : (4ll << (2 * retries));
        cap = std::max<long long>(32, std::min<long long>(wave * mult, 1ll << 28));
        shcap = std::max<long long>(32, std::min<long long>((tight ? wave : cap) * shadows_per_hit, 1ll << 28));
    };

retry:
    CK(cudaMemsetAsync(s->accum, 0, n_pix * sizeof(float4), st));
    CK(cudaMemsetAsync(s->counters, 0, DT_MAX_PIPES * DT_CNT_COUNT * sizeof(int), st));
    if (total == 0) {                                   // this rank owns no strip of the image (more ranks than strips): an empty share, not an error
        if (stats) *stats = S;
        return DT_OK;
    }
    // ---- sync-free wave loop: every ray of the frame fits one batch and the ray tree has a known depth bound, so all
    // waves are enqueued back to back (wave sizes live in device memory), shadow(k) runs on a second stream while
    // closest(k+1) / shade(k+1) proceed, and the frame is dealt to several such pipelines.  One host sync per frame.
    const bool bounded = !(pt && dc.russian_roulette) && s->dev.max_recursion_depth <= 16;
    if (!primary_only && !defer_mode && bounded && total <= (long long)wave_max && !host_loop) {
        const int n_waves = s->dev.max_recursion_depth + 1;
        int NP = s->n_pipes_env > 0 ? s->n_pipes_env : (int)1;
        NP = (int)std::min<long long>(std::min(NP, DT_MAX_PIPES), std::max<long long>(1, my_tiles));
        size_t ev_i = 0;
        auto ev = [&]() -> cudaEvent_t { if (ev_i >= s->ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); s->ev_pool.push_back(e); } return s->ev_pool[ev_i++]; };
        std::vector<cudaEvent_t> tg, tc, th, ts, tsort;      // (start, stop) pairs per stage
        auto timed = [&](std::vector<cudaEvent_t>& v, cudaStream_t q, auto&& launch) {
            cudaEvent_t a = ev(), b = ev();
            const unsigned int fl = s->capturing ? cudaEventRecordExternal : cudaEventRecordDefault;      // inside a capture: real event-record nodes
            cudaEventRecordWithFlags(a, q, fl); launch(); cudaEventRecordWithFlags(b, q, fl); v.push_back(a); v.push_back(b);
        };
        DtWaveParams wps[DT_MAX_PIPES]; int n0[DT_MAX_PIPES];
        for (int p = 0; p < NP; p++) {
            DtPipe& pp = s->pipes[p];
            wps[p] = wp;
            wps[p].tile_world = P.tile_world * NP;
            wps[p].tile_rank = P.tile_rank + p * P.tile_world;
            const long long tiles_p = dt_rank_tile_count(wp.tiles_x, wp.tiles_y, wps[p].tile_rank, wps[p].tile_world);
            wps[p].per_sample = tiles_p * 32;
            const long long total_p = wps[p].per_sample * dc.spp;
            n0[p] = (int)total_p;
            long long cap, shcap;
            queue_caps(total_p, cap, shcap);
            if ((rc = ensure_queues(s, pp, (int)cap, (int)shcap, false))) return rc;
        }
        // The whole frame -- generate, 7 x (closest, sort, shade, shadow, advance) on two streams, counter read-back -- is
        // captured ONCE into a CUDA graph and replayed while camera, seed, resolution and queue allocations stay the same
        // (the kernel parameters are baked into the graph): one cudaGraphLaunch per frame instead of ~50 launches and
        // ~150 event operations.  DT_GRAPH=0 enqueues directly (A/B).
        uint32_t n_launches = 0, n_closest = 0;
        auto enqueue_frame = [&]() -> int {
        CK(cudaEventRecord(s->ev_fork, st));
            for (int p = 0; p < NP; p++) {
                DtPipe& pp = s->pipes[p];
                if (p > 0) CK(cudaStreamWaitEvent(pp.A, s->ev_fork, 0));
                CK(cudaStreamWaitEvent(pp.B, s->ev_fork, 0));
                timed(tg, pp.A, [&] { if (n0[p] > 0) k_generate<<<(n0[p] + 255) / 256, 256, 0, pp.A>>>(dc, wps[p], pp.q[0], 0, 0, n0[p], s->accum); });
                s->h_counters[DT_MAX_PIPES * DT_CNT_COUNT + p] = n0[p];          // pinned scratch past the readback area
                CK(cudaMemcpyAsync(pp.counters + DT_CNT_CUR, s->h_counters + DT_MAX_PIPES * DT_CNT_COUNT + p, sizeof(int), cudaMemcpyHostToDevice, pp.A));
                n_launches++;
            }
            // Schedule per pipe (A = high-priority stream, B = low):
            //   A: closest(k) -> sort(k) -> shade(k) [fills shadow queue k % 3] -> advance(k) -> closest(k+1) ...
            //   B: shadow(k) is RELEASED BY advance(k), i.e. together with closest(k+1), and enqueued after it.  Both kernels are
            //      persistent and each can fill the GPU alone; released together, the high-priority closest(k+1) (the frame's
            //      critical path) gets the SM slots first and shadow(k) takes what it frees while it drains, then drains itself
            //      under shade(k+1) / closest(k+2).  With three shadow queues shadow(k) only has to finish before advance(k+2)
            //      recycles its queue.  Measured on config 2 (profiles/r1e_ab_shadow_order.log): 4.00 ms against 4.18 ms for
            //      DT_SHADOW_ORDER=0 (release shadow(k) right after shade(k), where it grabs every SM before closest(k+1) exists).
            //      Enforcing "closest first" strictly (a programmatic launch event fired once all closest blocks are resident)
            //      was slower, 4.3 ms, and a graph replay does not keep the order between its branches (4.1-4.2 ms), which is
            //      why enqueueing the launches is the default and the graph the option.
            auto launch_shadow = [&](DtPipe& pp, int k) {
                const int q = k % 3, cur = k & 1;
                int* c = pp.counters;
                timed(ts, pp.B, [&] { launch_traverse<true>(s, pp.q[cur], pp.sq[q], c + dt_cnt_shadow(q), 0, pp.shadow_capacity, c + dt_cnt_fetch_b(q), s->accum, pp.B, s->shadow_spare); });
                cudaEventRecord(pp.ev_shadow[q], pp.B);
                n_launches++;
            };
            for (int k = 0; k < n_waves; k++) {
                const int q = k % 3, cur = k & 1;
                for (int p = 0; p < NP; p++) {
                    DtPipe& pp = s->pipes[p];
                    int* c = pp.counters;
                    DtShadowQueue& sq = pp.sq[q];
                    timed(tc, pp.A, [&] { launch_traverse<false>(s, pp.q[cur], sq, c + DT_CNT_CUR, 0, pp.capacity, c + DT_CNT_FETCH_A, s->accum, pp.A); });
                    if (s->shadow_order && k >= 1) { CK(cudaStreamWaitEvent(pp.B, pp.ev_shade[(k - 1) % 3], 0)); launch_shadow(pp, k - 1); }
                    if (k >= 3) CK(cudaStreamWaitEvent(pp.A, pp.ev_shadow[q], 0));           // shadow(k-3) must have drained this queue
                    if (do_sort) timed(tsort, pp.A, [&] { n_launches += launch_sort(s, pp, pp.q[cur], c + DT_CNT_CUR, 0, pp.A, true, sort_spatial); });
                    timed(th, pp.A, [&] {
                        DtShadeCounters sc = {c + DT_CNT_NEXT, c + dt_cnt_shadow(q), c + DT_CNT_OVERFLOW, reinterpret_cast<unsigned long long*>(c + DT_CNT_SHADOW_DEAD), reinterpret_cast<unsigned long long*>(c + DT_CNT_CLOSEST_DEAD)};
                        (s->sphere_normal_maps ? k_shade<true> : k_shade<false>)<<<s->grid_shade, 128, 0, pp.A>>>(s->dev, dc, pp.q[cur], pp.miss[cur], c + DT_CNT_CUR, 0, do_sort ? pp.sort_perm : nullptr, pp.q[1 - cur], pp.miss[1 - cur], pp.capacity,
                                                                sq, pp.shadow_capacity, sc, s->accum); });
                    if (!s->shadow_order) { CK(cudaEventRecord(pp.ev_shade[q], pp.A)); CK(cudaStreamWaitEvent(pp.B, pp.ev_shade[q], 0)); launch_shadow(pp, k); }
                    if (k >= 2) CK(cudaStreamWaitEvent(pp.A, pp.ev_shadow[(k + 1) % 3], 0));  // shadow(k-2): its queue is recycled for wave k+1
                    k_wave_advance<<<1, 1, 0, pp.A>>>(c, (k + 1) % 3, pp.capacity, pp.shadow_capacity);
                    if (s->shadow_order) CK(cudaEventRecord(pp.ev_shade[q], pp.A));
                    n_launches += 3; n_closest++;
                }
            }
            for (int p = 0; p < NP; p++) {
                DtPipe& pp = s->pipes[p];
                if (s->shadow_order) { CK(cudaStreamWaitEvent(pp.B, pp.ev_shade[(n_waves - 1) % 3], 0)); launch_shadow(pp, n_waves - 1); }
                for (int q = 0; q < 3 && q < n_waves; q++) CK(cudaStreamWaitEvent(pp.A, pp.ev_shadow[q], 0));
                if (p > 0) { CK(cudaEventRecord(pp.ev_done, pp.A)); CK(cudaStreamWaitEvent(st, pp.ev_done, 0)); }
            }
            CK(cudaMemcpyAsync(s->h_counters, s->counters, DT_MAX_PIPES * DT_CNT_COUNT * sizeof(int), cudaMemcpyDeviceToHost, st));
    return DT_OK;
        };
        std::string key;
        if (frame_graph) {
            key.append((const char*)&dc, sizeof dc); key.append((const char*)wps, sizeof(DtWaveParams) * NP); key.append((const char*)n0, sizeof(int) * NP);
            const int misc[7] = {NP, n_waves, (do_sort ? 1 : 0) + 2 * sort_spatial + 64 * s->sort_refine, s->trav_mode, s->refill_threshold, (int)s->dev.n_shapes, s->shadow_order * 16 + s->shadow_spare};
            key.append((const char*)misc, sizeof misc);
            const void* ptrs[2] = {s->accum, s->counters};
            key.append((const char*)ptrs, sizeof ptrs);
            for (int p = 0; p < NP; p++) { const void* qp[3] = {s->pipes[p].q[0].o_time, s->pipes[p].sq[0].o_time, s->pipes[p].sort_perm}; key.append((const char*)qp, sizeof qp); }
        }
        if (frame_graph && s->frame_graph && key == s->frame_key) {
            tg = s->g_tg; tc = s->g_tc; th = s->g_th; ts = s->g_ts; tsort = s->g_tsort;
            n_launches = s->g_launches; n_closest = s->g_closest;
            CK(cudaGraphLaunch(s->frame_graph, st));
        } else if (frame_graph) {
            if (s->frame_graph) { cudaGraphExecDestroy(s->frame_graph); s->frame_graph = nullptr; s->frame_key.clear(); }
            while (s->ev_pool.size() < (size_t)(2 * NP * (1 + 4 * n_waves))) { cudaEvent_t e; CK(cudaEventCreate(&e)); s->ev_pool.push_back(e); }   // none created inside the capture
            s->capturing = true;
            cudaGraph_t g = nullptr;
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            const int erc = enqueue_frame();
            const cudaError_t ce = cudaStreamEndCapture(st, &g);
            s->capturing = false;
            if (erc) { if (g) cudaGraphDestroy(g); return erc; }
            if (ce != cudaSuccess || !g) { g_err = std::string("CUDA graph capture of the frame failed: ") + cudaGetErrorString(ce); return DT_ERR_CUDA; }
            const cudaError_t ie = cudaGraphInstantiate(&s->frame_graph, g, 0);
            cudaGraphDestroy(g);
            if (ie != cudaSuccess) { s->frame_graph = nullptr; g_err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie); return DT_ERR_CUDA; }
            s->frame_key = key;
            s->g_tg = tg; s->g_tc = tc; s->g_th = th; s->g_ts = ts; s->g_tsort = tsort; s->g_launches = n_launches; s->g_closest = n_closest;
            CK(cudaGraphLaunch(s->frame_graph, st));
        } else {
            if ((rc = enqueue_frame())) return rc;
        }
        S.kernel_launches += n_launches; S.launches_traverse_closest += n_closest;
        if (s->debug_timing) fprintf(stderr, "[dt] enqueue of %d pipes x %d waves took %.3f ms on the host\n", NP, n_waves,
                                     1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t_host0).count());
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        auto sum = [&](std::vector<cudaEvent_t>& v) { float t = 0.f; for (size_t i = 0; i + 1 < v.size(); i += 2) { float ms = 0.f; cudaEventElapsedTime(&ms, v[i], v[i + 1]); t += ms; } return t; };
        S.ms_generate = sum(tg); S.ms_traverse_closest = sum(tc); S.ms_shade = sum(th); S.ms_traverse_shadow = sum(ts); S.ms_sort = sum(tsort);
        S.waves = (uint32_t)n_waves;
        if (s->debug_timing >= 2 && !tg.empty()) {                                         // per-launch intervals relative to the frame's first launch
            auto dump = [&](const char* name, std::vector<cudaEvent_t>& v) {
                for (size_t i = 0; i + 1 < v.size(); i += 2) {
                    float a = 0.f, b = 0.f; cudaEventElapsedTime(&a, tg[0], v[i]); cudaEventElapsedTime(&b, tg[0], v[i + 1]);
                    fprintf(stderr, "[dt-tl] %-8s %2zu  %8.3f .. %8.3f  (%.3f ms)\n", name, i / 2, a, b, b - a);
                }
            };
            dump("generate", tg); dump("closest", tc); dump("sort", tsort); dump("shade", th); dump("shadow", ts);
        }
        bool overflow = false;
        unsigned long long tot_c = 0, tot_s = 0, tot_dead = 0;
        for (int p = 0; p < NP; p++) {
            const int* hc = s->h_counters + p * DT_CNT_COUNT;
            if (hc[DT_CNT_OVERFLOW] != 0) overflow = true;
            unsigned long long c8 = 0, s8 = 0, dead8 = 0;
            memcpy(&c8, hc + DT_CNT_TOT_CLOSEST, 8); memcpy(&s8, hc + DT_CNT_TOT_SHADOW, 8); memcpy(&dead8, hc + DT_CNT_SHADOW_DEAD, 8);
            unsigned long long cdead8 = 0;
            memcpy(&cdead8, hc + DT_CNT_CLOSEST_DEAD, 8);
            tot_c += c8 - std::min(c8, cdead8); tot_s += s8; tot_dead += dead8;
            for (int k = std::max(0, n_waves - 2); k < n_waves; k++)                          // the last two waves' queues were not recycled
                tot_s += (unsigned long long)std::min(hc[dt_cnt_shadow(k % 3)], s->pipes[p].shadow_capacity);
        }
        if (overflow) {
            if (retries >= 6 || wave_max <= 4096) { g_err = "wavefront queue overflow (ray-tree fan-out too large even for small waves)"; return DT_ERR_OVERFLOW; }
            retries++;
            wave_max = std::max(4096, (wave_max / 4 + 31) & ~31);      // smaller batches -> falls back to the synchronised loop
            goto retry;
        }
        S.rays_closest = tot_c + (uint64_t)(count_valid_pixels(P, W, H, dc.row_limit) * dc.spp);
        S.rays_shadow = tot_s - std::min(tot_s, tot_dead);
        S.retries = retries;
        if (stats) *stats = S;
        return DT_OK;
    }
    // ---- device-resident wave loop: unbounded depth (Russian roulette), deferred NEE, frames of several batches
    if (!primary_only && s->dev_loop && retries == 0 && !host_loop) {
        long long cap, shcap;
        queue_caps(wave_max, cap, shcap);
        bool overflow = false;
        if ((rc = render_devloop(s, dc, wp, total, wave_max, (int)cap, (int)shcap, shadows_per_hit, defer_mode, do_sort, sort_spatial, S, &overflow))) return rc;
        if (overflow) {                                   // retried in the host-synchronised loop with smaller waves
            retries++;
            wave_max = std::max(4096, (wave_max / 4 + 31) & ~31);
            goto retry;
        }
        S.rays_closest += (uint64_t)(count_valid_pixels(P, W, H, dc.row_limit) * dc.spp);
        S.retries = retries;
        if (stats) *stats = S;
        return DT_OK;
    }
    {
        DtPipe& pp = s->pipes[0];
        {
            long long cap, shcap;
            queue_caps(wave_max, cap, shcap);
            if (primary_only) { cap = wave_max; shcap = 32; }
            if ((rc = ensure_queues(s, pp, (int)cap, (int)shcap, defer_mode))) return rc;
        }
        int* c = pp.counters;
        DtShadowQueue& sq = pp.sq[0];
        long long next_primary = 0;
        int count = 0, cur = 0;
        int prev_shadow = 0;
        unsigned long long dead_total = 0, dead_prev_wave = 0;      // shadow-queue entries that are not traced (cumulative counter; those of the previous wave)
        unsigned long long cdead_total = 0, cdead_in_wave = 0;     // the same for the closest-hit queue: untraced entries sitting in the current wave
        bool overflow = false;
        S.rays_closest = 0; S.rays_shadow = 0; S.waves = 0; S.kernel_launches = 0; S.launches_traverse_closest = 0;
        S.ms_generate = S.ms_traverse_closest = S.ms_traverse_shadow = S.ms_shade = 0.f;
        while (next_primary < total || count > 0) {
            int n_new = 0;
            if (count < wave_max && next_primary < total) {
                n_new = (int)std::min<long long>(wave_max - count, total - next_primary);
                s->t_gen.start(st);
                k_generate<<<(n_new + 255) / 256, 256, 0, st>>>(dc, wp, pp.q[cur], count, next_primary, n_new, s->accum);
                s->t_gen.stop(st);
                S.kernel_launches++;
                count += n_new; next_primary += n_new;
            }
            CK(cudaMemsetAsync(c + DT_CNT_NEXT, 0, sizeof(int), st));
            CK(cudaMemsetAsync(c + DT_CNT_FETCH_A, 0, 2 * sizeof(int), st));
            s->t_closest.start(st);
            launch_traverse<false>(s, pp.q[cur], sq, nullptr, count, pp.capacity, c + DT_CNT_FETCH_A, s->accum);
            s->t_closest.stop(st);
            S.kernel_launches++; S.launches_traverse_closest++;
            if (primary_only) { CK(cudaStreamSynchronize(st)); S.ms_traverse_closest += s->t_closest.take(); S.ms_generate += s->t_gen.take(); S.waves++; break; }
            bool shadow_timed = false;
            if (defer_mode && prev_shadow > 0) {
                k_filter_deferred<<<(prev_shadow + 255) / 256, 256, 0, st>>>(s->dev, sq, prev_shadow, pp.q[cur]);
                s->t_shadow.start(st);
                launch_traverse<true>(s, pp.q[cur], sq, nullptr, prev_shadow, pp.shadow_capacity, c + DT_CNT_FETCH_B, s->accum);
                s->t_shadow.stop(st);
                shadow_timed = true;
                S.kernel_launches += 2;
            }
            CK(cudaMemsetAsync(c + DT_CNT_SHADOW, 0, sizeof(int), st));
            if (do_sort) { s->t_sort.start(st); S.kernel_launches += launch_sort(s, pp, pp.q[cur], nullptr, count, st, true, sort_spatial); s->t_sort.stop(st); }
            s->t_shade.start(st);
            {
                DtShadeCounters sc = {c + DT_CNT_NEXT, c + DT_CNT_SHADOW, c + DT_CNT_OVERFLOW, reinterpret_cast<unsigned long long*>(c + DT_CNT_SHADOW_DEAD), reinterpret_cast<unsigned long long*>(c + DT_CNT_CLOSEST_DEAD)};
                (s->sphere_normal_maps ? k_shade<true> : k_shade<false>)<<<(count + 127) / 128, 128, 0, st>>>(s->dev, dc, pp.q[cur], pp.miss[cur], nullptr, count, do_sort ? pp.sort_perm : nullptr, pp.q[1 - cur], pp.miss[1 - cur], pp.capacity,
                                                             sq, pp.shadow_capacity, sc, s->accum);
            }
            s->t_shade.stop(st);
            S.kernel_launches++;
            if (!defer_mode) {
                s->t_shadow.start(st);
                launch_traverse<true>(s, pp.q[cur], sq, c + DT_CNT_SHADOW, 0, pp.shadow_capacity, c + DT_CNT_FETCH_B, s->accum);
                s->t_shadow.stop(st);
                shadow_timed = true;
                S.kernel_launches++;
            }
            CK(cudaMemcpyAsync(s->h_counters, c, DT_CNT_COUNT * sizeof(int), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaGetLastError());
            S.ms_generate += s->t_gen.take();
            const float ms_c = s->t_closest.take(), ms_s = shadow_timed ? s->t_shadow.take() : 0.f;
            S.ms_traverse_closest += ms_c;
            S.ms_shade += s->t_shade.take();
            S.ms_sort += s->t_sort.take();
            S.ms_traverse_shadow += ms_s;
            S.waves++;
            unsigned long long dead_now = 0;
            memcpy(&dead_now, s->h_counters + DT_CNT_SHADOW_DEAD, 8);
            const unsigned long long dead_this_wave = dead_now - dead_total;     // emitted (and marked) by this wave's shade pass
            dead_total = dead_now;
            if (count >= wave_max / 2) {                                   // a "bulk" wave (dt_stats)
                S.bulk_waves++;
                S.ms_bulk_closest += ms_c; S.rays_bulk_closest += (uint64_t)count - std::min<uint64_t>((uint64_t)count, cdead_in_wave);
                const unsigned long long entries = (unsigned long long)(defer_mode ? prev_shadow : std::min(s->h_counters[DT_CNT_SHADOW], pp.shadow_capacity));
                const unsigned long long dead_traced_now = defer_mode ? dead_prev_wave : dead_this_wave;      // the pass timed in ms_s traced these entries
                S.ms_bulk_shadow += ms_s; S.rays_bulk_shadow += (uint64_t)(entries - std::min(entries, dead_traced_now));
            }
            dead_prev_wave = dead_this_wave;
            {
                unsigned long long cdead_now = 0;
                memcpy(&cdead_now, s->h_counters + DT_CNT_CLOSEST_DEAD, 8);
                cdead_in_wave = cdead_now - cdead_total;               // marked by this wave's shade pass: they sit in the NEXT wave's queue
                cdead_total = cdead_now;
            }
            if (s->debug_timing >= 2) fprintf(stderr, "[dt-tl] wave %u: %d closest rays, %d shadow rays emitted (%llu of them not traced) | closest %.3f sort %.3f shade %.3f shadow %.3f ms (cumulative)\n", S.waves - 1, count,
                                              s->h_counters[DT_CNT_SHADOW], dead_this_wave, S.ms_traverse_closest, S.ms_sort, S.ms_shade, S.ms_traverse_shadow);
            if (s->h_counters[DT_CNT_OVERFLOW] != 0) { overflow = true; break; }
            const int next_count = s->h_counters[DT_CNT_NEXT];
            const int shadow_count = std::min(s->h_counters[DT_CNT_SHADOW], pp.shadow_capacity);
            S.rays_closest += (uint64_t)next_count;
            S.rays_shadow += (uint64_t)shadow_count;
            prev_shadow = defer_mode ? shadow_count : 0;
            count = next_count;
            cur = 1 - cur;
        }
        if (!overflow && defer_mode && prev_shadow > 0) {
            CK(cudaMemsetAsync(c + DT_CNT_FETCH_B, 0, sizeof(int), st));
            launch_traverse<true>(s, pp.q[cur], sq, nullptr, prev_shadow, pp.shadow_capacity, c + DT_CNT_FETCH_B, s->accum);
            S.kernel_launches++;
        }
        if (overflow) {
            if (retries >= 6 || wave_max <= 4096) { g_err = "wavefront queue overflow (ray-tree fan-out too large even for small waves)"; return DT_ERR_OVERFLOW; }
            retries++;
            wave_max = std::max(4096, (wave_max / 4 + 31) & ~31);
            CK(cudaStreamSynchronize(st));
            goto retry;
        }
        S.rays_closest += (uint64_t)(count_valid_pixels(P, W, H, dc.row_limit) * (primary_only ? 1 : dc.spp));
        S.rays_shadow -= std::min<uint64_t>(S.rays_shadow, dead_total);
        S.rays_closest -= std::min<uint64_t>(S.rays_closest, cdead_total);
    }
    S.retries = retries;
    if (stats) *stats = S;
    return DT_OK;
}

int finish_core(dt_scene* s, const dt_camera_desc* cam, const float* hdr_dev, int flags, uint8_t* ldr_host, dt_stats* S) {
    const int W = cam->width, H = cam->height;
    const size_t n_pix = (size_t)W * H;
    cudaStream_t st = s->stream;
    if (cam->has_tonemapper && !(flags & DT_FLAG_SKIP_TONEMAP)) {
        s->t_tm.start(st);
        uint32_t l = 0;
        int rc = tonemap_device(s, hdr_dev, W, H, cam->tm_key, cam->tm_burn, cam->tm_saturation, cam->tm_gamma, s->ldr, &l);
        if (rc) return rc;
        s->t_tm.stop(st);
        if (S) S->kernel_launches += l;
    }
    if (ldr_host) CK(cudaMemcpyAsync(ldr_host, s->ldr, n_pix * 3, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (S) S->ms_tonemap = s->t_tm.take();
    return DT_OK;
}

}  // namespace

extern "C" {

int dt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int dt_gpu_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) { g_err = "no CUDA device available; this library has no CPU fallback"; return DT_ERR_NO_DEVICE; }
    if (device < 0 || device >= n) { g_err = "device index out of range"; return DT_ERR_INVALID; }
    g_device = device;
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_err = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return DT_ERR_CUDA; }
    return n;
}

int dt_scene_create(const dt_scene_desc* desc, dt_scene** out) { return dt_scene_create_opts(desc, nullptr, out); }

static int scene_create_impl(const dt_scene_desc* desc, const dt_scene_options* opts, dt_scene** out);
int dt_scene_create_opts(const dt_scene_desc* desc, const dt_scene_options* opts, dt_scene** out) {
    const int rc = scene_create_impl(desc, opts, out);
    dt_resident_trees_clear();           // device copies of dt_bvh2_build trees this scene did not consume (dt_build.cu)
    return rc;
}
static int scene_create_impl(const dt_scene_desc* desc, const dt_scene_options* opts, dt_scene** out) {
    if (!desc || !out) { g_err = "null argument"; return DT_ERR_INVALID; }
    *out = nullptr;
    DtHostScene hs;
    std::string err;
    // meshes of at least this many faces get their BLAS from the GPU flattener (dt_flatten_gpu.cu); smaller ones are not worth its launches
    int gpu_min_faces = 32768;
    if (opts && opts->gpu_flatten_min_faces != 0) gpu_min_faces = opts->gpu_flatten_min_faces < 0 ? 0x7FFFFFFF : opts->gpu_flatten_min_faces;
    if (!dt_flatten_scene(desc, hs, err, gpu_min_faces)) { g_err = "scene description rejected: " + err; return DT_ERR_INVALID; }
    int rc = ensure_device();
    if (rc) return rc;
    dt_scene* s = new dt_scene();
    s->device = tl_device >= 0 ? tl_device : g_device;
    memset(&s->dev, 0, sizeof s->dev);
    auto fail = [&](int code) { dt_scene_destroy(s); return code; };
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, s->device) != cudaSuccess) { g_err = "cudaGetDeviceProperties failed"; return fail(DT_ERR_CUDA); }
    s->num_sms = prop.multiProcessorCount;
    // the closest-hit -> shade chain is the frame's critical path: its stream gets the high priority, the any-hit (shadow) stream the low one
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (const char* e = getenv("DT_STREAM_PRIO")) { if (atoi(e) == 0) prio_lo = prio_hi = 0; }
    if (cudaStreamCreateWithPriority(&s->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess || cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming) != cudaSuccess) { g_err = "cudaStreamCreate failed"; return fail(DT_ERR_CUDA); }
    for (int p = 0; p < DT_MAX_PIPES; p++) {
        DtPipe& pp = s->pipes[p];
        bool ok = true;
        if (p == 0) pp.A = s->stream; else ok = cudaStreamCreateWithPriority(&pp.A, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
        ok = ok && cudaStreamCreateWithPriority(&pp.B, cudaStreamNonBlocking, prio_lo) == cudaSuccess && cudaEventCreateWithFlags(&pp.ev_done, cudaEventDisableTiming) == cudaSuccess;
        for (int k = 0; k < 3 && ok; k++) ok = cudaEventCreateWithFlags(&pp.ev_shade[k], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&pp.ev_shadow[k], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { g_err = "cudaStreamCreate / cudaEventCreate failed"; return fail(DT_ERR_CUDA); }
    }
    DtSceneDev& D = s->dev;
    {
        const uint4* p = nullptr;
        if ((rc = upload<uint4>(s->allocs, (const uint4*)hs.tlas_nodes.data(), hs.tlas_nodes.size() * 5, &p))) return fail(rc);
        D.tlas_nodes = p;
    }
    if ((rc = upload<int32_t>(s->allocs, hs.tlas_prims.data(), hs.tlas_prims.size(), &D.tlas_prims))) return fail(rc);
    if ((rc = upload<DtFaceDev>(s->allocs, hs.faces.data(), hs.faces.size(), &D.faces))) return fail(rc);
    if ((rc = upload<float>(s->allocs, hs.verts.data(), hs.verts.size(), &D.verts))) return fail(rc);
    {
        // BLAS arrays: the host-flattened meshes first, then one slice per GPU-flattened mesh (in mesh order)
        size_t gpu_faces = 0;
        for (int mi : hs.gpu_meshes) gpu_faces += (size_t)desc->meshes[mi].n_faces;
        const size_t host_nodes = hs.blas_nodes.size(), host_prims = hs.tris.size() / 3;
        const size_t node_cap = host_nodes + gpu_faces, n_prims = host_prims + gpu_faces;      // a BVH8 over n single-face leaves has fewer than n nodes
        if (n_prims > 0x7FFFFFF0ull || node_cap > 0x7FFFFFF0ull) { g_err = "scene too large for 32-bit primitive indices"; return fail(DT_ERR_INVALID); }
        const uint4* nodes = nullptr;
        if ((rc = upload<uint4>(s->allocs, (const uint4*)hs.blas_nodes.data(), host_nodes * 5, &nodes, (node_cap - host_nodes) * 5))) return fail(rc);
        if ((rc = upload<float4>(s->allocs, hs.tris.data(), hs.tris.size(), &D.tris, gpu_faces * 3))) return fail(rc);
        if ((rc = upload<float4>(s->allocs, hs.leaf_boxes.data(), hs.leaf_boxes.size(), &D.leaf_boxes, gpu_faces * 2))) return fail(rc);
        if ((rc = upload<uint32_t>(s->allocs, hs.face_prim.data(), std::min(hs.face_prim.size(), hs.faces.size()), &D.face_prim, hs.faces.size() - std::min(hs.face_prim.size(), hs.faces.size())))) return fail(rc);
        size_t n_nodes = host_nodes, prim_off = host_prims;
        int blas_depth = hs.blas_depth;
        for (int mi : hs.gpu_meshes) {
            const dt_mesh& m = desc->meshes[mi];
            DtMeshDev& md = hs.meshes[(size_t)mi];
            uint32_t made = 0; int depth = 0;
            md.node_root = (uint32_t)n_nodes;
            if (!dt_flatten_mesh_gpu(m, D.faces + md.face_base, D.verts + (size_t)md.vert_base * 3, (DtNode8*)nodes, (uint32_t)n_nodes, (uint32_t)(node_cap - n_nodes),
                                     (float4*)D.tris, (float4*)D.leaf_boxes, (uint32_t*)D.face_prim + md.face_base, (uint32_t)prim_off, &made, &depth, err)) {
                g_err = "scene description rejected: " + err; return fail(DT_ERR_INVALID);
            }
            n_nodes += made; prim_off += (size_t)m.n_faces;
            blas_depth = std::max(blas_depth, depth);
        }
        if (hs.tlas_depth + blas_depth + 4 > DT_STACK_SIZE) {
            g_err = "scene description rejected: BVH too deep for the traversal stack (" + std::to_string(hs.tlas_depth + blas_depth + 4) + " > " + std::to_string(DT_STACK_SIZE) + ")";
            return fail(DT_ERR_INVALID);
        }
        if (node_cap - n_nodes > (1u << 16)) {                     // give the unused tail of the node allocation back
            const uint4* exact = nullptr;
            if ((rc = upload<uint4>(s->allocs, (const uint4*)nullptr, 0, &exact, n_nodes * 5))) return fail(rc);
            if (cudaMemcpy((void*)exact, nodes, n_nodes * sizeof(DtNode8), cudaMemcpyDeviceToDevice) != cudaSuccess) { g_err = "device copy of the BVH8 nodes failed"; return fail(DT_ERR_CUDA); }
            for (auto it = s->allocs.begin(); it != s->allocs.end(); ++it) if (*it == (void*)nodes) { s->allocs.erase(it); break; }
            cudaFree((void*)nodes);
            nodes = exact;
        }
        D.blas_nodes = nodes;
        s->n_blas_nodes = n_nodes; s->n_prims = n_prims; s->n_faces = hs.faces.size();
    }
    if ((rc = upload<DtShapeDev>(s->allocs, hs.shapes.data(), hs.shapes.size(), &D.shapes))) return fail(rc);
    if ((rc = upload<DtMeshDev>(s->allocs, hs.meshes.data(), hs.meshes.size(), &D.meshes))) return fail(rc);
    if ((rc = upload<float>(s->allocs, hs.uvs.data(), hs.uvs.size(), &D.uvs))) return fail(rc);
    if ((rc = upload<float>(s->allocs, hs.vnormals.data(), hs.vnormals.size(), &D.vnormals))) return fail(rc);
    if ((rc = upload<dt_material>(s->allocs, desc->materials, (size_t)desc->n_materials, &D.materials))) return fail(rc);
    if ((rc = upload<dt_brdf>(s->allocs, desc->brdfs, (size_t)desc->n_brdfs, &D.brdfs))) return fail(rc);
    if ((rc = upload<dt_point_light>(s->allocs, desc->point_lights, (size_t)desc->n_point_lights, &D.point_lights))) return fail(rc);
    if ((rc = upload<dt_area_light>(s->allocs, desc->area_lights, (size_t)desc->n_area_lights, &D.area_lights))) return fail(rc);
    if ((rc = upload<dt_directional_light>(s->allocs, desc->directional_lights, (size_t)desc->n_directional_lights, &D.directional_lights))) return fail(rc);
    if ((rc = upload<dt_spot_light>(s->allocs, desc->spot_lights, (size_t)desc->n_spot_lights, &D.spot_lights))) return fail(rc);
    if ((rc = upload<dt_env_light>(s->allocs, desc->env_lights, (size_t)desc->n_env_lights, &D.env_lights))) return fail(rc);
    if ((rc = upload<dt_mesh_light>(s->allocs, desc->mesh_lights, (size_t)desc->n_mesh_lights, &D.mesh_lights))) return fail(rc);
    if ((rc = upload<dt_texture>(s->allocs, desc->textures, (size_t)desc->n_textures, &D.textures))) return fail(rc);
    if ((rc = upload<DtImageDev>(s->allocs, hs.images.data(), hs.images.size(), &D.images))) return fail(rc);
    if ((rc = upload<uint8_t>(s->allocs, hs.image_u8.data(), hs.image_u8.size(), &D.image_u8, 16))) return fail(rc);
    if ((rc = upload<float>(s->allocs, hs.image_f32.data(), hs.image_f32.size(), &D.image_f32, 16))) return fail(rc);
    D.n_shapes = desc->n_shapes; D.n_mesh_shapes = desc->n_mesh_shapes; D.n_materials = desc->n_materials;
    D.n_point_lights = desc->n_point_lights; D.n_area_lights = desc->n_area_lights; D.n_directional_lights = desc->n_directional_lights;
    D.n_spot_lights = desc->n_spot_lights; D.n_env_lights = desc->n_env_lights; D.n_mesh_lights = desc->n_mesh_lights;
    D.bg_texture = desc->bg_texture; D.max_recursion_depth = desc->max_recursion_depth;
    D.tlas_direct = 0;
    D.one_bits = 0x3F800000u;
    if (hs.tlas_nodes.size() == 1 && hs.tlas_prims.size() <= 4 && hs.tlas_nodes[0].imask == 0 && !getenv("DT_NO_TLAS_DIRECT")) D.tlas_direct = (int)hs.tlas_prims.size();
    memcpy(D.background_color, desc->background_color, 12);
    D.shadow_ray_epsilon = desc->shadow_ray_epsilon;
    memcpy(D.ambient_light, desc->ambient_light, 12);
    for (int a = 0; a < 3; a++) {
        const float ext = hs.world_max[a] - hs.world_min[a];
        D.sort_min[a] = hs.world_min[a];
        D.sort_scale[a] = (ext > 0.f && ext < 1e30f) ? 1.0f / ext : 0.f;
    }
    s->n_triangles = hs.n_triangles;
    s->lights_shadowed = desc->n_point_lights + desc->n_area_lights + desc->n_directional_lights + desc->n_spot_lights + desc->n_mesh_lights;
    s->has_env = desc->n_env_lights > 0;
    for (int i = 0; i < desc->n_shapes; i++) if (desc->shapes[i].kind == DT_SHAPE_SPHERE && desc->shapes[i].tex_normal >= 0) s->sphere_normal_maps = true;
    bool any_diel = false, any_refl = false;
    for (int i = 0; i < desc->n_materials; i++) {
        if (desc->materials[i].type == DT_MAT_DIELECTRIC) any_diel = true;
        if (desc->materials[i].type == DT_MAT_MIRROR || desc->materials[i].type == DT_MAT_CONDUCTOR) any_refl = true;
    }
    s->fanout_hint = desc->max_recursion_depth > 0 ? (any_diel ? 2 : (any_refl ? 1 : 0)) : 0;

    if (cudaMalloc(&s->counters, DT_MAX_PIPES * DT_CNT_COUNT * sizeof(int)) != cudaSuccess || cudaMallocHost(&s->h_counters, (DT_MAX_PIPES * DT_CNT_COUNT + DT_MAX_PIPES) * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&s->tm_logsum, sizeof(double)) != cudaSuccess || cudaMalloc(&s->tm_hist, 256 * sizeof(unsigned int)) != cudaSuccess ||
        cudaMalloc(&s->tm_rank, sizeof(unsigned long long)) != cudaSuccess || cudaMalloc(&s->tm_prefix, sizeof(uint32_t)) != cudaSuccess) {
        g_err = "cudaMalloc of control buffers failed"; return fail(DT_ERR_CUDA);
    }
    for (int p = 0; p < DT_MAX_PIPES; p++) s->pipes[p].counters = s->counters + p * DT_CNT_COUNT;
    for (Timer* t : {&s->t_total, &s->t_gen, &s->t_closest, &s->t_shadow, &s->t_shade, &s->t_sort, &s->t_resolve, &s->t_tm}) if (t->init()) return fail(DT_ERR_CUDA);
    // traversal variant (A/B measurement): 0 static if-if, 1 static while-while, 2 dynamic if-if, 3 dynamic while-while
    if (const char* e = getenv("DT_TRAVERSE_MODE")) s->trav_mode = std::min(3, std::max(0, atoi(e)));
    if (const char* e = getenv("DT_SYNC_WAVES")) s->sync_waves = atoi(e);
    if (const char* e = getenv("DT_SORT")) s->sort_mode = std::min(2, std::max(0, atoi(e)));
    if (const char* e = getenv("DT_SORT_SPATIAL")) s->sort_spatial = std::min(DT_SSORT_MAX_AXIS_BITS, std::max(-1, atoi(e)));
    if (const char* e = getenv("DT_SORT_REFINE")) s->sort_refine = atoi(e) != 0;
    if (const char* e = getenv("DT_SHADOW_ORDER")) s->shadow_order = atoi(e);
    if (const char* e = getenv("DT_SHADOW_SPARE")) s->shadow_spare = std::max(0, atoi(e));
    if (const char* e = getenv("DT_GRAPH")) s->use_graph = atoi(e);
    if (const char* e = getenv("DT_DEBUG_TIMING")) s->debug_timing = atoi(e);
    if (const char* e = getenv("DT_DEVLOOP")) s->dev_loop = atoi(e);
    if (const char* e = getenv("DT_TAIL_THRESHOLD")) s->tail_threshold = std::max(-1, atoi(e));
    if (const char* e = getenv("DT_PIPES")) s->n_pipes_env = std::min(DT_MAX_PIPES, std::max(0, atoi(e)));
    if (const char* e = getenv("DT_REFILL_THRESHOLD")) s->refill_threshold = std::min(32, std::max(1, atoi(e)));
    {
        int bps = 0;
        const void* fns[4][2] = {{(const void*)k_traverse<false, false>, (const void*)k_traverse<true, false>},
                                 {(const void*)k_traverse<false, true>, (const void*)k_traverse<true, true>},
                                 {(const void*)k_traverse_dyn<false, false>, (const void*)k_traverse_dyn<true, false>},
                                 {(const void*)k_traverse_dyn<false, true>, (const void*)k_traverse_dyn<true, true>}};
        for (int m = 0; m < 4; m++) for (int a = 0; a < 2; a++) {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, fns[m][a], 128, 0);
            if (const char* e = getenv(a ? "DT_BPS_ANY" : "DT_BPS_CLOSEST")) bps = std::min(bps, std::max(1, atoi(e)));      // experiment knob: blocks per SM
            s->grid_trav[m][a] = s->num_sms * std::max(1, bps);
        }
    }
    { int bps = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_shade<false>, 128, 0); s->grid_shade = s->num_sms * std::max(1, bps); }
    if (cudaDeviceSynchronize() != cudaSuccess) { g_err = "device sync after upload failed"; return fail(DT_ERR_CUDA); }
    *out = s;
    return DT_OK;
}

void dt_scene_destroy(dt_scene* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    dt_frame_release(s);
    s->free_loop();
    s->free_tail();
    s->free_queues();
    for (void* p : s->allocs) cudaFree(p);
    if (s->counters) cudaFree(s->counters);
    if (s->h_counters) cudaFreeHost(s->h_counters);
    if (s->accum) cudaFree(s->accum);
    if (s->hdr) cudaFree(s->hdr);
    if (s->ldr) cudaFree(s->ldr);
    if (s->tm_logsum) cudaFree(s->tm_logsum);
    if (s->tm_hist) cudaFree(s->tm_hist);
    if (s->tm_rank) cudaFree(s->tm_rank);
    if (s->tm_prefix) cudaFree(s->tm_prefix);
    for (Timer* t : {&s->t_total, &s->t_gen, &s->t_closest, &s->t_shadow, &s->t_shade, &s->t_sort, &s->t_resolve, &s->t_tm}) t->destroy();
    if (s->frame_graph) cudaGraphExecDestroy(s->frame_graph);
    for (cudaEvent_t e : s->ev_pool) cudaEventDestroy(e);
    for (int p = 0; p < DT_MAX_PIPES; p++) {
        DtPipe& pp = s->pipes[p];
        for (int k = 0; k < 3; k++) { if (pp.ev_shade[k]) cudaEventDestroy(pp.ev_shade[k]); if (pp.ev_shadow[k]) cudaEventDestroy(pp.ev_shadow[k]); }
        if (pp.ev_done) cudaEventDestroy(pp.ev_done);
        if (pp.B) cudaStreamDestroy(pp.B);
        if (p > 0 && pp.A) cudaStreamDestroy(pp.A);
    }
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int dt_render_device(dt_scene* s, const dt_camera_desc* cam, const dt_render_params* params, float** hdr_dev, dt_stats* stats) {
    if (!s || !hdr_dev) { g_err = "null argument"; return DT_ERR_INVALID; }
    dt_stats S; memset(&S, 0, sizeof S);
    int rc = render_core(s, cam, params, &S, false);
    if (rc) return rc;
    cudaStream_t st = s->stream;
    const int n_pix = cam->width * cam->height;
    const int spp = cam->samples_per_pixel < 1 ? 1 : cam->samples_per_pixel;
    const bool peer_frame = params && (params->flags & DT_FLAG_PEER_FRAME);
    if (peer_frame && s->peer_hdr && (s->peer_w != cam->width || s->peer_h != cam->height)) { g_err = "imported peer frame has a different resolution"; return DT_ERR_INVALID; }
    s->t_resolve.start(st);
    if (peer_frame) {
        float* dst_hdr = s->hdr; uint8_t* dst_ldr = s->ldr;
        if (s->peer_hdr) {
            const bool want_hdr = cam->has_tonemapper || (params->flags & DT_FLAG_PEER_HDR);
            dst_hdr = want_hdr ? s->peer_hdr : nullptr;                   // the destination tonemaps the whole frame from radiance ...
            dst_ldr = cam->has_tonemapper ? nullptr : s->peer_ldr;        // ... or only needs the clamped bytes (main.cpp:118-125)
        }
        const int tiles_x = (cam->width + 7) / 8, tiles_y = (cam->height + 3) / 4;
        const int world = params->tile_world < 1 ? 1 : params->tile_world;
        const long long my_tiles = dt_rank_tile_count(tiles_x, tiles_y, params->tile_rank, world);
        const long long threads = my_tiles / DT_TILE_GROUP * 32;     // one warp per strip
        if (threads > 0) k_resolve_tiles<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(s->accum, cam->width, cam->height, tiles_x, tiles_y, my_tiles, params->tile_rank, world, spp, dst_hdr, dst_ldr, s->counters);
    } else {
        k_resolve<<<(n_pix + 255) / 256, 256, 0, st>>>(s->accum, n_pix, spp, s->hdr, s->ldr, s->counters);
    }
    s->t_resolve.stop(st);
    S.kernel_launches++;
    s->t_total.stop(st);
    CK(cudaMemcpyAsync(s->h_counters, s->counters, DT_CNT_COUNT * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    S.nan_pixels = (uint64_t)s->h_counters[DT_CNT_NAN];
    S.ms_resolve = s->t_resolve.take();
    S.ms_total = s->t_total.take();
    *hdr_dev = s->hdr;
    if (stats) *stats = S;
    return DT_OK;
}

int dt_finish_device(dt_scene* s, const dt_camera_desc* cam, const float* hdr_dev, uint8_t* ldr_rgb, dt_stats* stats) {
    if (!s || !cam || !hdr_dev || !ldr_rgb) { g_err = "null argument"; return DT_ERR_INVALID; }
    int rc = check_cam(cam);
    if (rc) return rc;
    CK(cudaSetDevice(s->device));
    const size_t n_pix = (size_t)cam->width * cam->height;
    if ((rc = ensure_outputs(s, n_pix))) return rc;
    dt_stats S; memset(&S, 0, sizeof S);
    cudaStream_t st = s->stream;
    s->t_total.start(st);
    // copy into our own hdr buffer if it is a foreign pointer, then finish
    if (hdr_dev != s->hdr) CK(cudaMemcpyAsync(s->hdr, hdr_dev, n_pix * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (!cam->has_tonemapper) {
        // clamp from s->hdr
        k_clamp_hdr<<<((int)n_pix + 255) / 256, 256, 0, st>>>(s->hdr, (int)n_pix, s->ldr);
        S.kernel_launches++;
    }
    rc = finish_core(s, cam, s->hdr, 0, ldr_rgb, &S);
    if (rc) return rc;
    s->t_total.stop(st);
    CK(cudaStreamSynchronize(st));
    S.ms_total = s->t_total.take();
    if (stats) *stats = S;
    return DT_OK;
}

int dt_frame_export(dt_scene* s, int32_t width, int32_t height, dt_frame_handle* out) {
    if (!s || !out || width <= 0 || height <= 0) { g_err = "bad argument"; return DT_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    int rc = ensure_outputs(s, (size_t)width * height);
    if (rc) return rc;
    memset(out, 0, sizeof *out);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "dt_frame_handle stores cudaIpcMemHandle_t as 64 bytes");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, s->hdr)); memcpy(out->hdr, &h, 64);
    CK(cudaIpcGetMemHandle(&h, s->ldr)); memcpy(out->ldr, &h, 64);
    out->width = width; out->height = height;
    s->frame_exported = true;
    return DT_OK;
}

int dt_frame_release(dt_scene* s) {
    if (!s) return DT_OK;
    cudaSetDevice(s->device);
    if (s->peer_is_ipc) {
        if (s->peer_hdr) cudaIpcCloseMemHandle(s->peer_hdr);
        if (s->peer_ldr) cudaIpcCloseMemHandle(s->peer_ldr);
    }
    s->peer_hdr = nullptr; s->peer_ldr = nullptr; s->peer_w = s->peer_h = 0; s->peer_is_ipc = false;
    return DT_OK;
}

int dt_frame_import(dt_scene* s, const dt_frame_handle* in) {
    if (!s || !in) { g_err = "null argument"; return DT_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    dt_frame_release(s);
    cudaIpcMemHandle_t h;
    memcpy(&h, in->hdr, 64);
    CK(cudaIpcOpenMemHandle((void**)&s->peer_hdr, h, cudaIpcMemLazyEnablePeerAccess));
    memcpy(&h, in->ldr, 64);
    CK(cudaIpcOpenMemHandle((void**)&s->peer_ldr, h, cudaIpcMemLazyEnablePeerAccess));
    s->peer_w = in->width; s->peer_h = in->height; s->peer_is_ipc = true;
    return DT_OK;
}

int dt_frame_finish(dt_scene* s, const dt_camera_desc* cam, uint8_t* ldr_rgb, dt_stats* stats) {
    if (!s || !cam) { g_err = "null argument"; return DT_ERR_INVALID; }
    int rc = check_cam(cam);
    if (rc) return rc;
    CK(cudaSetDevice(s->device));
    if (s->accum_pix < (size_t)cam->width * cam->height) { g_err = "no frame of this size has been rendered or exported"; return DT_ERR_INVALID; }
    dt_stats S; memset(&S, 0, sizeof S);
    cudaStream_t st = s->stream;
    s->t_total.start(st);
    rc = finish_core(s, cam, s->hdr, 0, ldr_rgb, &S);
    if (rc) return rc;
    s->t_total.stop(st);
    CK(cudaStreamSynchronize(st));
    S.ms_total = s->t_total.take();
    if (stats) *stats = S;
    return DT_OK;
}

int dt_render(dt_scene* s, const dt_camera_desc* cam, const dt_render_params* params, uint8_t* ldr_rgb, float* hdr_rgb, dt_stats* stats) {
    if (!s || !ldr_rgb) { g_err = "null argument"; return DT_ERR_INVALID; }
    float* hdr_dev = nullptr;
    dt_stats S; memset(&S, 0, sizeof S);
    int rc = dt_render_device(s, cam, params, &hdr_dev, &S);
    if (rc) return rc;
    const size_t n_pix = (size_t)cam->width * cam->height;
    cudaStream_t st = s->stream;
    const int flags = params ? params->flags : 0;
    if (hdr_rgb) CK(cudaMemcpyAsync(hdr_rgb, hdr_dev, n_pix * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    rc = finish_core(s, cam, hdr_dev, flags, ldr_rgb, &S);
    if (rc) return rc;
    S.ms_total += S.ms_tonemap;
    if (stats) *stats = S;
    return DT_OK;
}


// ------------------------------------------------------------------ one process, several GPUs
// The reference's main() is one process that splits rows over its threads (main.cpp:38-39,164-185).  dt_multi does the same
// over the GPUs of the box: the scene is replicated, one host thread per GPU renders that GPU's strips (tile_rank = device,
// tile_world = n), every resolve kernel stores its strips straight into device 0's frame through peer access over NVLink
// (the DT_FLAG_PEER_FRAME path, with plain peer pointers instead of CUDA IPC mappings), the thread join is the barrier, and
// device 0 tonemaps / copies the complete frame out.
struct dt_multi {
    std::vector<dt_scene*> scenes;
};

int dt_multi_create(const dt_scene_desc* desc, int n_devices, dt_multi** out) {
    if (!desc || !out) { g_err = "null argument"; return DT_ERR_INVALID; }
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { g_err = "no CUDA device available; this library has no CPU fallback"; return DT_ERR_NO_DEVICE; }
    const int n = n_devices <= 0 ? count : n_devices;
    if (n > count) { g_err = "dt_multi_create: " + std::to_string(n) + " devices requested, " + std::to_string(count) + " visible"; return DT_ERR_INVALID; }
    dt_multi* m = new dt_multi();
    m->scenes.assign((size_t)n, nullptr);
    std::vector<int> rcs((size_t)n, DT_OK);
    std::vector<std::string> errs((size_t)n);
    std::vector<std::thread> th;
    for (int d = 0; d < n; d++) th.emplace_back([&, d] {
        tl_device = d;
        rcs[(size_t)d] = dt_scene_create_opts(desc, nullptr, &m->scenes[(size_t)d]);
        if (rcs[(size_t)d] == DT_OK && d > 0) {
            int can = 0;
            cudaSetDevice(d);
            if (cudaDeviceCanAccessPeer(&can, d, 0) != cudaSuccess || !can) { rcs[(size_t)d] = DT_ERR_UNSUPPORTED; g_err = "device " + std::to_string(d) + " has no peer access to device 0"; }
            else { const cudaError_t e = cudaDeviceEnablePeerAccess(0, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { rcs[(size_t)d] = DT_ERR_CUDA; g_err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); } cudaGetLastError(); }
        }
        errs[(size_t)d] = g_err;
        tl_device = -1;
    });
    for (auto& t : th) t.join();
    for (int d = 0; d < n; d++) if (rcs[(size_t)d] != DT_OK) {
        g_err = "dt_multi_create, device " + std::to_string(d) + ": " + errs[(size_t)d];
        const int rc = rcs[(size_t)d];
        dt_multi_destroy(m);
        return rc;
    }
    *out = m;
    return DT_OK;
}

void dt_multi_destroy(dt_multi* m) {
    if (!m) return;
    for (dt_scene* s : m->scenes) if (s) { s->peer_hdr = nullptr; s->peer_ldr = nullptr; dt_scene_destroy(s); }
    delete m;
}

int dt_multi_device_count(const dt_multi* m) { return m ? (int)m->scenes.size() : 0; }

int dt_multi_render(dt_multi* m, const dt_camera_desc* cam, const dt_render_params* params, uint8_t* ldr_rgb, float* hdr_rgb, dt_stats* stats) {
    if (!m || !ldr_rgb) { g_err = "null argument"; return DT_ERR_INVALID; }
    int rc = check_cam(cam);
    if (rc) return rc;
    const int n = (int)m->scenes.size();
    dt_scene* s0 = m->scenes[0];
    CK(cudaSetDevice(s0->device));
    const size_t n_pix = (size_t)cam->width * cam->height;
    if ((rc = ensure_outputs(s0, n_pix))) return rc;
    for (int d = 1; d < n; d++) { dt_scene* s = m->scenes[(size_t)d]; s->peer_hdr = s0->hdr; s->peer_ldr = s0->ldr; s->peer_w = cam->width; s->peer_h = cam->height; s->peer_is_ipc = false; }
    dt_render_params P; memset(&P, 0, sizeof P); P.seed = 1234;
    if (params) P = *params;
    P.tile_world = n;
    P.flags |= DT_FLAG_PEER_FRAME | (hdr_rgb ? DT_FLAG_PEER_HDR : 0);
    std::vector<int> rcs((size_t)n, DT_OK);
    std::vector<std::string> errs((size_t)n);
    std::vector<dt_stats> st((size_t)n);
    std::vector<std::thread> th;
    for (int d = 0; d < n; d++) th.emplace_back([&, d] {
        dt_render_params Pd = P; Pd.tile_rank = d;
        float* hdr_dev = nullptr;
        rcs[(size_t)d] = dt_render_device(m->scenes[(size_t)d], cam, &Pd, &hdr_dev, &st[(size_t)d]);
        errs[(size_t)d] = g_err;
    });
    for (auto& t : th) t.join();                                                   // = "all strips are in device 0's frame"
    for (int d = 0; d < n; d++) if (rcs[(size_t)d] != DT_OK) { g_err = "dt_multi_render, device " + std::to_string(d) + ": " + errs[(size_t)d]; return rcs[(size_t)d]; }
    dt_stats S = st[0];
    for (int d = 1; d < n; d++) {
        const dt_stats& t = st[(size_t)d];
        S.rays_closest += t.rays_closest; S.rays_shadow += t.rays_shadow; S.nan_pixels += t.nan_pixels; S.kernel_launches += t.kernel_launches;
        S.waves = std::max(S.waves, t.waves); S.ms_total = std::max(S.ms_total, t.ms_total); S.retries = std::max(S.retries, t.retries);
    }
    CK(cudaSetDevice(s0->device));
    if (hdr_rgb) CK(cudaMemcpyAsync(hdr_rgb, s0->hdr, n_pix * 3 * sizeof(float), cudaMemcpyDeviceToHost, s0->stream));
    dt_stats F; memset(&F, 0, sizeof F);
    if ((rc = finish_core(s0, cam, s0->hdr, P.flags, ldr_rgb, &F))) return rc;
    S.kernel_launches += F.kernel_launches; S.ms_tonemap = F.ms_tonemap; S.ms_total += F.ms_tonemap;
    if (stats) *stats = S;
    return DT_OK;
}

int dt_primary_hits(dt_scene* s, const dt_camera_desc* cam, int32_t* shape, int32_t* face, float* t) {
    if (!s || !cam || !shape || !face || !t) { g_err = "null argument"; return DT_ERR_INVALID; }
    { const int crc = check_cam(cam); if (crc) return crc; }
    dt_render_params P; memset(&P, 0, sizeof P); P.seed = 1234; P.tile_world = 1;
    const long long n_slots = dt_rank_tile_count((cam->width + 7) / 8, (cam->height + 3) / 4, 0, 1) * 32;      // ray slots of the frame (whole strips)
    if (n_slots > (1ll << 28)) { g_err = "image too large for dt_primary_hits"; return DT_ERR_INVALID; }
    P.max_wave_rays = (int)n_slots;
    dt_stats S;
    int rc = render_core(s, cam, &P, &S, true);
    if (rc) return rc;
    const int n_pix = cam->width * cam->height;
    int32_t *d_shape = nullptr, *d_face = nullptr; float* d_t = nullptr;
    if (cudaMalloc(&d_shape, (size_t)n_pix * 4) != cudaSuccess || cudaMalloc(&d_face, (size_t)n_pix * 4) != cudaSuccess || cudaMalloc(&d_t, (size_t)n_pix * 4) != cudaSuccess) {
        cudaFree(d_shape); cudaFree(d_face); cudaFree(d_t);
        g_err = "dt_primary_hits: cudaMalloc failed"; return DT_ERR_CUDA;
    }
    cudaStream_t st = s->stream;
    k_unpack_hits<<<((int)n_slots + 255) / 256, 256, 0, st>>>(s->pipes[0].q[0].hit0, s->pipes[0].q[0].hit_face, s->pipes[0].q[0].pixel, (int)n_slots, d_shape, d_face, d_t, 1);
    cudaMemcpyAsync(shape, d_shape, (size_t)n_pix * 4, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(face, d_face, (size_t)n_pix * 4, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(t, d_t, (size_t)n_pix * 4, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(d_shape); cudaFree(d_face); cudaFree(d_t);
    if (e != cudaSuccess) { g_err = std::string("dt_primary_hits: ") + cudaGetErrorString(e); return DT_ERR_CUDA; }
    return DT_OK;
}

static int trace_generic(dt_scene* s, const float* origins, const float* dirs, const float* tmax, int64_t n, bool any,
                         int32_t* shape, int32_t* face, float* t, uint8_t* occluded) {
    if (!s || !origins || !dirs || n < 0) { g_err = "bad argument"; return DT_ERR_INVALID; }
    if (n == 0) return DT_OK;
    if (n > (1ll << 28)) { g_err = "too many rays in one call"; return DT_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    cudaStream_t st = s->stream;
    float *d_o = nullptr, *d_d = nullptr, *d_tm = nullptr; float4 *o4 = nullptr, *d4 = nullptr, *hit0 = nullptr; int32_t* hface = nullptr;
    int32_t *d_shape = nullptr, *d_face = nullptr; float* d_t = nullptr; uint8_t* d_occ = nullptr; int* fetch = nullptr;
    std::vector<void*> tmp;
    auto A = [&](void** p, size_t b) -> int { CK(cudaMalloc(p, std::max<size_t>(b, 16))); tmp.push_back(*p); return DT_OK; };
    int rc = DT_OK;
    do {
        if ((rc = A((void**)&d_o, (size_t)n * 12)) || (rc = A((void**)&d_d, (size_t)n * 12)) || (rc = A((void**)&o4, (size_t)n * 16)) || (rc = A((void**)&d4, (size_t)n * 16)) || (rc = A((void**)&fetch, 16))) break;
        cudaMemcpyAsync(d_o, origins, (size_t)n * 12, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_d, dirs, (size_t)n * 12, cudaMemcpyHostToDevice, st);
        if (tmax) { if ((rc = A((void**)&d_tm, (size_t)n * 4))) break; cudaMemcpyAsync(d_tm, tmax, (size_t)n * 4, cudaMemcpyHostToDevice, st); }
        cudaMemsetAsync(fetch, 0, 16, st);
        k_pack_rays<<<((int)n + 255) / 256, 256, 0, st>>>(d_o, d_d, d_tm, (int)n, o4, d4);
        if (any) {
            if ((rc = A((void**)&d_occ, (size_t)n))) break;
            k_traverse_occluded<<<((int)n + 127) / 128, 128, 0, st>>>(s->dev, o4, d4, (int)n, d_occ);
            cudaMemcpyAsync(occluded, d_occ, (size_t)n, cudaMemcpyDeviceToHost, st);
        } else {
            if ((rc = A((void**)&hit0, (size_t)n * 16)) || (rc = A((void**)&hface, (size_t)n * 4)) || (rc = A((void**)&d_shape, (size_t)n * 4)) || (rc = A((void**)&d_face, (size_t)n * 4)) || (rc = A((void**)&d_t, (size_t)n * 4))) break;
            DtRayQueue q; memset(&q, 0, sizeof q);
            q.o_time = o4; q.d_tmax = d4; q.hit0 = hit0; q.hit_face = hface; q.pixel = nullptr;
            DtShadowQueue sq; memset(&sq, 0, sizeof sq);
            launch_traverse<false>(s, q, sq, nullptr, (int)n, (int)n, fetch, nullptr);
            k_unpack_hits<<<((int)n + 255) / 256, 256, 0, st>>>(hit0, hface, nullptr, (int)n, d_shape, d_face, d_t, 0);
            cudaMemcpyAsync(shape, d_shape, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
            cudaMemcpyAsync(face, d_face, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
            cudaMemcpyAsync(t, d_t, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
        }
    } while (0);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaError_t e2 = cudaGetLastError();
    for (void* p : tmp) cudaFree(p);
    if (rc) return rc;
    if (e != cudaSuccess || e2 != cudaSuccess) { g_err = std::string("trace: ") + cudaGetErrorString(e != cudaSuccess ? e : e2); return DT_ERR_CUDA; }
    return DT_OK;
}

int dt_trace_closest(dt_scene* s, const float* origins, const float* dirs, int64_t n, int32_t* shape, int32_t* face, float* t) {
    if (!shape || !face || !t) { g_err = "null output"; return DT_ERR_INVALID; }
    return trace_generic(s, origins, dirs, nullptr, n, false, shape, face, t, nullptr);
}
int dt_trace_occluded(dt_scene* s, const float* origins, const float* dirs, const float* tmax, int64_t n, uint8_t* occluded) {
    if (!occluded || !tmax) { g_err = "null argument"; return DT_ERR_INVALID; }
    return trace_generic(s, origins, dirs, tmax, n, true, nullptr, nullptr, nullptr, occluded);
}

int dt_tonemap(const float* hdr_rgb, int32_t width, int32_t height, float key, float burn, float saturation, float gamma, uint8_t* ldr_rgb) {
    if (!hdr_rgb || !ldr_rgb || width <= 0 || height <= 0) { g_err = "bad argument"; return DT_ERR_INVALID; }
    int rc = ensure_device();
    if (rc) return rc;
    // a throw-away scene-less context: allocate just what the tonemapper needs
    dt_scene tmp;
    tmp.device = g_device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, tmp.device));
    tmp.num_sms = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&tmp.stream, cudaStreamNonBlocking));
    const size_t n_pix = (size_t)width * height;
    float* d_hdr = nullptr; uint8_t* d_ldr = nullptr;
    rc = DT_OK;
    do {
        if (cudaMalloc(&d_hdr, n_pix * 12) != cudaSuccess || cudaMalloc(&d_ldr, n_pix * 3) != cudaSuccess || cudaMalloc(&tmp.tm_logsum, 8) != cudaSuccess ||
            cudaMalloc(&tmp.tm_hist, 1024) != cudaSuccess || cudaMalloc(&tmp.tm_rank, 8) != cudaSuccess || cudaMalloc(&tmp.tm_prefix, 4) != cudaSuccess) { g_err = "cudaMalloc failed"; rc = DT_ERR_CUDA; break; }
        cudaMemcpyAsync(d_hdr, hdr_rgb, n_pix * 12, cudaMemcpyHostToDevice, tmp.stream);
        uint32_t l = 0;
        rc = tonemap_device(&tmp, d_hdr, width, height, key, burn, saturation, gamma, d_ldr, &l);
        if (rc) break;
        cudaMemcpyAsync(ldr_rgb, d_ldr, n_pix * 3, cudaMemcpyDeviceToHost, tmp.stream);
        if (cudaStreamSynchronize(tmp.stream) != cudaSuccess) { g_err = "tonemap sync failed"; rc = DT_ERR_CUDA; }
    } while (0);
    cudaFree(d_hdr); cudaFree(d_ldr); cudaFree(tmp.tm_logsum); cudaFree(tmp.tm_hist); cudaFree(tmp.tm_rank); cudaFree(tmp.tm_prefix);
    cudaStreamDestroy(tmp.stream);
    return rc;
}

#ifdef DT_TRAV_STATS
void dt_debug_stats(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_dt_stats, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_dt_stats, z, sizeof z); }
}
#endif

#ifdef DT_TIMELINE
// debug build only: copy out (and reset) the per-warp timeline records; returns the number of records
int dt_debug_timeline(unsigned long long* out, int max_records) {
    cudaDeviceSynchronize();
    unsigned int n = 0;
    cudaMemcpyFromSymbol(&n, g_dt_tl_count, sizeof n);
    if (n > (unsigned)max_records) n = (unsigned)max_records;
    if (n > DT_TL_MAX) n = DT_TL_MAX;
    if (out && n) cudaMemcpyFromSymbol(out, g_dt_tl, (size_t)n * 32);
    unsigned int z = 0; cudaMemcpyToSymbol(g_dt_tl_count, &z, sizeof z);
    return (int)n;
}
void dt_debug_steps_hist(unsigned int* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_dt_steps_hist, sizeof(unsigned int) * 128);
    if (reset) { unsigned int z[128] = {0}; cudaMemcpyToSymbol(g_dt_steps_hist, z, sizeof z); }
}
#endif

int dt_scene_accel_checksum(dt_scene* s, uint64_t out[10]) {
    if (!s || !out) { g_err = "null argument"; return DT_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    CK(cudaDeviceSynchronize());
    std::string err;
    const void* ptr[4] = {s->dev.blas_nodes, s->dev.tris, s->dev.leaf_boxes, s->dev.face_prim};
    const size_t bytes[4] = {s->n_blas_nodes * sizeof(DtNode8), s->n_prims * 48, s->n_prims * 32, s->n_faces * 4};
    for (int k = 0; k < 4; k++) if (!dt_device_checksum(ptr[k], bytes[k], out + 2 * k, err)) { g_err = err; return DT_ERR_CUDA; }
    out[8] = s->n_blas_nodes; out[9] = s->n_prims;
    return DT_OK;
}

void* dt_scene_stream(dt_scene* s) { return s ? (void*)s->stream : nullptr; }

const char* dt_last_error(void) { return g_err.c_str(); }
const char* dt_version(void) { return "dorktracer-b200 0.1 (sm_100a, abi 1)"; }

}  // extern "C"

