// Wavefront kernels of the render hot path (sm_100a).  One wave =
//   k_generate (top-up with new camera samples) -> k_traverse<closest> -> k_shade (accumulate local terms,
//   emit shadow rays + child rays) -> k_traverse<shadow> (accumulate unoccluded light terms)
// replacing the per-pixel recursion of Raytracer::PerPixel / PerformShading (raytracer.cpp:38-134).
// The recursion is linear in child radiance, so every tree node carries its RGB weight W and adds
// W * Local(node) straight into the pixel accumulator (SURVEY.md 8a "Ray-tree semantics").
#pragma once
#include "dt_device.h"
#include "dt_math.cuh"
#include "dt_traverse.cuh"
#include "dt_shade.cuh"

struct DtCamDev {
    float position[3], gaze[3], up[3], right[3], q[3];
    float left, right_, bottom, top;
    int width, height, spp;
    float focus_distance, aperture_size;
    int path_tracing, importance_sampling, nee, russian_roulette;
    int jitter_aa;                 // DT_FLAG_JITTER_AA: keep the sub-pixel sample position (off = the reference's int truncation, main.cpp:83)
    int smooth_shading;            // DT_FLAG_SMOOTH_SHADING: interpolated vertex normals on meshes that carry them (not in the reference)
    int keep_weightless;           // DT_FLAG_KEEP_WEIGHTLESS_PATHS: shade hits whose path weight is exactly zero too (what the reference does)
    int row_limit;                 // rows y >= row_limit get no camera rays (DT_FLAG_REF_ROW_BANDS: the reference's 8 row bands leave the
                                   // bottom H mod 8 rows unrendered, main.cpp:38-39); height otherwise
};

// Tile ownership across ranks.  Every row of 8x4-pixel tiles is
// cut into STRIPS of DT_TILE_GROUP tiles (64x4 pixels; the last strip of a row may be shorter) and the strips are dealt out
// round-robin in row-major strip order, so that a rank's resolve kernel writes 192 contiguous bytes per row segment into the
// destination rank's frame over NVLink instead of 24.  j-th tile slot of a rank -> tile (tx, ty); false = dead slot.
#define DT_TILE_GROUP 8
__host__ __device__ __forceinline__ bool dt_rank_tile(long long j, int rank, int world, int tiles_x, int tiles_y, int& tx, int& ty) {
    const int strips_x = (tiles_x + DT_TILE_GROUP - 1) / DT_TILE_GROUP;
    const long long strip = (j / DT_TILE_GROUP) * world + rank;
    ty = (int)(strip / strips_x);
    tx = (int)(strip % strips_x) * DT_TILE_GROUP + (int)(j % DT_TILE_GROUP);
    return ty < tiles_y && tx < tiles_x;
}
// tile slots of a rank (whole strips; slots past the end of a short strip are dead)
__host__ __device__ __forceinline__ long long dt_rank_tile_count(int tiles_x, int tiles_y, int rank, int world) {
    const long long strips = (long long)((tiles_x + DT_TILE_GROUP - 1) / DT_TILE_GROUP) * tiles_y;
    return rank < strips ? ((strips - rank + world - 1) / world) * DT_TILE_GROUP : 0;
}

struct DtWaveParams {
    uint32_t seed_lo, seed_hi;
    int tile_rank, tile_world;
    int tiles_x, tiles_y;
    long long per_sample;          // primary slots per sample on this rank (tiles owned * 32)
};

// counters[] layout (device ints)
enum { DT_CNT_NEXT = 0, DT_CNT_SHADOW = 1, DT_CNT_FETCH_A = 2, DT_CNT_FETCH_B = 3, DT_CNT_NAN = 4, DT_CNT_OVERFLOW = 5, DT_CNT_DEFER = 6,
       DT_CNT_CUR = 7,            // rays in the current wave (device-resident loop)
       DT_CNT_SHADOW2 = 8, DT_CNT_FETCH_B2 = 9,   // second shadow queue (shadow(k) overlaps closest(k+1))
       DT_CNT_SHADOW3 = 14, DT_CNT_FETCH_B3 = 15, // third shadow queue (shadow(k) may run until advance(k+2))
       DT_CNT_TOT_CLOSEST = 10, DT_CNT_TOT_SHADOW = 12,   // 64-bit totals (two ints each)
       // device-resident wave loop (k_loop_begin / k_loop_end / k_tail)
       DT_CNT_PREV_SHADOW = 16,   // deferred-NEE shadow rays of the previous wave, traced after this wave's closest-hit pass
       DT_CNT_GEN_N = 17, DT_CNT_GEN_BASE = 18,           // top-up of this wave: new camera samples, first queue slot
       DT_CNT_WAVES = 19,         // waves that held any work
       DT_CNT_NEXT_PRIMARY = 20,  // 64-bit: camera samples generated so far
       DT_CNT_GEN_K0 = 22,        // 64-bit: first camera-sample index of this wave's top-up
       DT_CNT_ITERS = 24,         // executions of the WHILE body (two waves each)
       DT_CNT_TAIL_WAVES = 25,    // longest block-local wave chain of k_tail
       DT_CNT_TAIL_RAYS = 26,     // rays handed to k_tail
       DT_CNT_SHADOW_DEAD = 28,   // 64-bit: shadow-queue entries of this frame that are not traced (zero contribution, see dt_shade_ray)
       DT_CNT_CLOSEST_DEAD = 30,  // 64-bit: closest-hit queue entries of this frame that are not traced (GI children of zero weight)
       DT_CNT_COUNT = 32 };
struct DtShadeCounters { int* next; int* shadow; int* overflow; unsigned long long* shadow_dead; unsigned long long* closest_dead; };

#define DT_DEAD_PIXEL 0xFFFFFFFFu
// Russian roulette never ends a path whose throughput is NaN or stays >= 1 (`probTest > maxThroughput` is false, raytracer.cpp:141-146):
// in a closed part of a scene the reference recurses until its stack overflows (a crash after ~10^4 bounces).  The wavefront loop
// would spin forever instead, so a path is cut this many bounces below depth 0 -- further than the reference can get.
#define DT_RR_MAX_BOUNCES 32768

__device__ __forceinline__ int dt_agg_inc(int* counter) {
    const unsigned active = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(active) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(active));
    base = __shfl_sync(active, base, leader);
    return base + __popc(active & ((1u << lane) - 1u));
}

// k slots for every active lane with ONE atomic per warp (k is warp-uniform).  The warp's block of slots is laid out
// item-major: item j of the lane with rank r is slot  first + j * stride  (first = return value, stride = active lanes), so the
// lanes' stores of one item are consecutive and the j-th items of neighbouring lanes (e.g. their shadow rays towards the same
// light) stay neighbours in the queue.
__device__ __forceinline__ int dt_agg_reserve(int* counter, int k, int& stride) {
    const unsigned active = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(active) - 1;
    stride = __popc(active);
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, stride * k);
    base = __shfl_sync(active, base, leader);
    return base + __popc(active & ((1u << lane) - 1u));
}

__device__ __forceinline__ void dt_accum(float4* accum, uint32_t pix, v3 c) {
    float* p = reinterpret_cast<float*>(accum + pix);
    atomicAdd(p + 0, c.x); atomicAdd(p + 1, c.y); atomicAdd(p + 2, c.z);
}

// ------------------------------------------------------------------ generate
// Camera::GetImagePlanePosition (camera.cpp:74-80) + Raytracer::GenerateRay (raytracer.cpp:661-699) + the
// stratified-sample / Gaussian-weight part of renderThreadMain (main.cpp:59-96).
// (jx, jy) = position inside the pixel: 0.5 (the pixel centre, camera.cpp:76-77) unless DT_FLAG_JITTER_AA
__device__ inline void dt_camera_ray(const DtCamDev& cam, int i, int j, float jx, float jy, DtRng& rng, v3& o, v3& d, float& mb_time) {
    float su = (float)((i + (double)jx) * (double)(cam.right_ - cam.left) / cam.width);
    float sv = (float)((j + (double)jy) * (double)(cam.top - cam.bottom) / cam.height);
    v3 ipp = vadd(vadd(F3(cam.q), vscale(F3(cam.right), su)), vscale(F3(cam.up), -sv));
    o = F3(cam.position);
    if (cam.aperture_size > 0.0001) {
        v3 ap = o;
        float first01 = -1.0f + 2.0f * rng01(rng);
        ap = vadd(ap, vscale(F3(cam.up), (first01 * cam.aperture_size * 0.5f)));
        float second01 = -1.0f + 2.0f * rng01(rng);
        ap = vadd(ap, vscale(F3(cam.right), (second01 * cam.aperture_size * 0.5f)));
        v3 dir = vunit(vsub(o, ipp));
        float tFd = cam.focus_distance / vdot(dir, F3(cam.gaze));
        v3 bent = vadd(o, vscale(dir, tFd));
        d = vunit(vsub(bent, ap));
        o = ap;
    } else {
        d = vunit(vsub(ipp, o));
    }
    mb_time = rng01(rng);
}

__device__ __forceinline__ void dt_generate_one(const DtCamDev& cam, const DtWaveParams& wp, const DtRayQueue& q, const int slot, const long long k, float4* accum) {
    const int s = (int)(k / wp.per_sample);
    const long long rem = k % wp.per_sample;
    int tx, ty;
    const bool live = dt_rank_tile(rem >> 5, wp.tile_rank, wp.tile_world, wp.tiles_x, wp.tiles_y, tx, ty);
    const int lane = (int)(rem & 31);
    const int x = tx * 8 + (lane & 7);
    const int y = ty * 4 + (lane >> 3);
    if (!live || x >= cam.width || y >= cam.row_limit) {
        q.pixel[slot] = DT_DEAD_PIXEL;
        return;
    }
    const uint32_t pix = (uint32_t)(x + y * cam.width);
    DtRng rng; rng.key = dt_hash(dt_hash(pix, (uint32_t)s) ^ wp.seed_lo, wp.seed_hi); rng.ctr = 0;
    int px = x, py = y;
    float w = 1.0f;
    float cam_jx = 0.5f, cam_jy = 0.5f;
    if (cam.spp > 1) {
        const int nRows = (int)sqrt((double)cam.spp), nCols = nRows;
        const int st = s % (nRows * nCols);
        const int row = st / nCols, col = st % nCols;
        float psi1 = rng01(rng), psi2 = rng01(rng);
        float sx = (col + psi1) / nCols;
        float sy = (row + psi2) / nRows;
        // A non-square sample count: the reference fills nRows^2 entries of `samples` but iterates samplesPerPixel of them
        // (main.cpp:47,63-81); the rest keep the zeros the vector was created with -- rays through the pixel with the Gaussian weight
        // of its corner.  DT_FLAG_JITTER_AA gives them a stratum instead.
        if (!cam.jitter_aa && s >= nRows * nCols) { sx = 0.0f; sy = 0.0f; }
        if (cam.jitter_aa) {
            // opt-in (DT_FLAG_JITTER_AA): the sample position keeps its sub-pixel jitter.  The reference means to do
            // this but RenderPixel(int,int,...) truncates it away (main.cpp:83), so parity mode leaves it off.
            cam_jx = sx; cam_jy = sy;
        }
        if (!cam.jitter_aa) {
            px = (int)(sx + x);      // RenderPixel(int,int,...) truncates the float sample position (main.cpp:83)
            py = (int)(sy + y);
        }
        const float sigma = 1.0f / 6.0f;                     // gaussian.h:3-21, main.cpp:52
        const float sigmaSqr = sigma * sigma;
        const float c1 = (float)(1.0f / (2.0f * DT_PI * sigmaSqr));
        float xd = sx - 0.5f, yd = sy - 0.5f;
        float exponent = (float)(-0.5 * (double)((xd * xd + yd * yd) / sigmaSqr));
        w = c1 * expf(exponent);
        atomicAdd(reinterpret_cast<float*>(accum + pix) + 3, w);
    }
    v3 o, d; float mb;
    dt_camera_ray(cam, px, py, cam_jx, cam_jy, rng, o, d, mb);
    q.o_time[slot] = make_float4(o.x, o.y, o.z, mb);
    q.d_tmax[slot] = make_float4(d.x, d.y, d.z, CUDART_INF_F);
    q.pixel[slot] = pix;
    q.weight_n[slot] = make_float4(w, w, w, 1.0f);
    q.thr_beer[slot] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    // misc.y of a camera ray (no parent material): the sample's truncated pixel coordinate relative to the accumulation pixel.
    // (int)(sx + x) is x except when the float sum rounds up to x + 1; the reference's background-texture lookup uses that
    // coordinate (raytracer.cpp:49-55 with main.cpp:83), so it travels with the ray.
    q.misc[slot] = make_int4(0x7FFFFFFF /* set by shade from the scene */, (px - x) | ((py - y) << 1), (int)dt_hash(rng.key, 0x9E37u), DT_FLAG_PRIMARY);
}

__global__ void k_generate(DtCamDev cam, DtWaveParams wp, DtRayQueue q, int base, long long k0, int n, float4* accum) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    dt_generate_one(cam, wp, q, base + idx, k0 + idx, accum);
}
// device-resident wave loop: the top-up size / slot / first sample index of this wave were written by k_loop_begin
__global__ void k_generate_dev(DtCamDev cam, DtWaveParams wp, DtRayQueue q, const int* c, float4* accum) {
    const int n = c[DT_CNT_GEN_N], base = c[DT_CNT_GEN_BASE];
    const long long k0 = *reinterpret_cast<const long long*>(c + DT_CNT_GEN_K0);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x)
        dt_generate_one(cam, wp, q, base + idx, k0 + idx, accum);
}

// ------------------------------------------------------------------ traversal (persistent warps)
__device__ __forceinline__ void dt_store_closest(const DtRayQueue& q, int i, const DtHit& h) {
    q.hit0[i] = make_float4(h.t, h.beta, h.gamma, __int_as_float(h.shape));
    q.hit_face[i] = h.face;
}
__device__ __forceinline__ void dt_store_shadow(const DtShadowQueue& sq, int i, const DtHit& h, float4* accum) {
    if (h.shape < 0) {
        const float4 c = sq.contrib_pix[i];
        dt_accum(accum, (uint32_t)__float_as_int(c.w), V(c.x, c.y, c.z));
    }
}

#ifndef DT_TRAV_MINBLOCKS
#define DT_TRAV_MINBLOCKS 7       // closest-hit kernel: resident blocks per SM the register allocator must allow (72 registers)
#endif
#ifndef DT_TRAV_MINBLOCKS_ANY
#define DT_TRAV_MINBLOCKS_ANY 8   // any-hit kernel (less state: no barycentrics, no tie-breaking): 64 registers, no spills
#endif
// Static variant: a warp fetches 32 consecutive rays and runs them to completion.
template <bool ANY, bool WW>
__global__ void __launch_bounds__(128, ANY ? DT_TRAV_MINBLOCKS_ANY : DT_TRAV_MINBLOCKS) k_traverse(DtSceneDev S, DtRayQueue q, DtShadowQueue sq, const int* n_ptr, int n_fixed, int n_cap, int* fetch_counter, float4* accum) {
    const int n = min(n_ptr ? *n_ptr : n_fixed, n_cap);        // a device-side counter may have run past the queue capacity (overflow): never index beyond it
    const int lane = threadIdx.x & 31;
    DT_DECLARE_STACK(stack);
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(fetch_counter, 32);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        const int i = base + lane;
        if (i < n) {
            if (!ANY && q.pixel && q.pixel[i] == DT_DEAD_PIXEL) continue;
            const float4 o = ANY ? sq.o_time[i] : q.o_time[i];
            const float4 d = ANY ? sq.d_tmax[i] : q.d_tmax[i];
            if (ANY && d.w < 0.0f) continue;                            // zero contribution: not traced (dt_shade_ray)
            DtTrav T;
            dt_trav_init<ANY>(T, S, V(o.x, o.y, o.z), V(d.x, d.y, d.z), o.w, ANY ? d.w : CUDART_INF_F);
            while (!dt_trav_step<ANY, WW>(T, stack, S, (ANY ? sq.o_time : q.o_time) + i, (ANY ? sq.d_tmax : q.d_tmax) + i)) {}
            if (ANY) dt_store_shadow(sq, i, T.best, accum); else dt_store_closest(q, i, T.best);
        }
    }
}

// Dynamic variant: every lane is a resumable traversal state machine; when fewer than `refill_threshold`
// lanes of the warp still hold a ray, the idle lanes fetch new rays with ONE warp-aggregated atomic
// (Aila-Laine style persistent threads).  Keeps SIMT lanes busy on incoherent secondary / shadow rays.
#ifdef DT_TIMELINE
// debug build only: per-warp (kind|n, t_start, t_drained, t_exit) records in nanoseconds of %globaltimer
#define DT_TL_MAX (1 << 20)
__device__ unsigned long long g_dt_tl[DT_TL_MAX * 4];
__device__ unsigned int g_dt_tl_count;
__device__ unsigned int g_dt_steps_hist[2][64];      // [ANY][min(63, steps / 8)] per-ray state-machine steps
__device__ __forceinline__ unsigned long long dt_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif

#ifndef DT_TRI_BATCH
#define DT_TRI_BATCH 0            // A/B knob: > 0 = lanes with pending primitives park until this many lanes hold some (dt_trav_step_nodes / _prims)
#endif
template <bool ANY, bool WW>
__global__ void __launch_bounds__(128, ANY ? DT_TRAV_MINBLOCKS_ANY : DT_TRAV_MINBLOCKS) k_traverse_dyn(DtSceneDev S, DtRayQueue q, DtShadowQueue sq, const int* n_ptr, int n_fixed, int n_cap, int* fetch_counter,
                                                      float4* accum, int refill_threshold) {
    const int n = min(n_ptr ? *n_ptr : n_fixed, n_cap);        // a device-side counter may have run past the queue capacity (overflow): never index beyond it
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lanes_lt = (1u << lane) - 1u;
    DtTrav T;
    DT_DECLARE_STACK(stack);
    int ray = -1;
    bool drained = false;
#ifdef DT_TIMELINE
    const unsigned long long tl_t0 = dt_now(); unsigned long long tl_td = 0; int tl_steps = 0;
#endif
    for (;;) {
        // One fetch round: the idle lanes take the next rays of the queue with ONE warp-aggregated atomic.
        auto fetch_round = [&](const unsigned idle) {
            const int leader = __ffs(idle) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(fetch_counter, __popc(idle));
            base = __shfl_sync(FULL, base, leader);
            if (ray < 0) {
                const int i = base + __popc(idle & lanes_lt);
                if (i < n && (ANY || !q.pixel || q.pixel[i] != DT_DEAD_PIXEL)) {
                    const float4 d = ANY ? sq.d_tmax[i] : q.d_tmax[i];
                    if (!ANY || !(d.w < 0.0f)) {                    // tmax < 0: a shadow-queue entry that is not traced (zero contribution, dt_shade_ray)
                        const float4 o = ANY ? sq.o_time[i] : q.o_time[i];
                        dt_trav_init<ANY>(T, S, V(o.x, o.y, o.z), V(d.x, d.y, d.z), o.w, ANY ? d.w : CUDART_INF_F);
                        ray = i;
                    }
                }
            }
            if (base + __popc(idle) >= n) {
                drained = true;
#ifdef DT_TIMELINE
                tl_td = dt_now();
#endif
            }
        };
        if (!drained) {
            const unsigned idle = __ballot_sync(FULL, ray < 0);
            if (idle) fetch_round(idle);
            if (ANY) {
                // untraced entries come in runs of 32 (a shading warp's shadow rays towards one light are neighbours in the queue):
                // lanes that drew one fetch again, up to three more rounds while at least 8 lanes are idle
                for (int round = 0; round < 3 && !drained; round++) {
                    const unsigned again = __ballot_sync(FULL, ray < 0);
                    if (__popc(again) < 8) break;
                    fetch_round(again);
                }
            }
        }
        unsigned act = __ballot_sync(FULL, ray >= 0);
        if (act == 0u) { if (drained) break; else continue; }
        for (;;) {
#if DT_TRI_BATCH > 0
            {
                const float4* ro = (ANY ? sq.o_time : q.o_time) + ray;
                const float4* rd = (ANY ? sq.d_tmax : q.d_tmax) + ray;
                bool fin = false;
                if (ray >= 0 && T.tg.y == 0u) fin = dt_trav_step_nodes<ANY>(T, stack, S, ro, rd);
                const bool has_prims = ray >= 0 && !fin && T.tg.y != 0u;
                const unsigned pend = __ballot_sync(FULL, has_prims);
                const unsigned adv = __ballot_sync(FULL, ray >= 0 && !fin && T.tg.y == 0u);
                if (pend != 0u && (__popc(pend) >= DT_TRI_BATCH || adv == 0u)) {
                    if (has_prims) fin = dt_trav_step_prims<ANY>(T, stack, S, ro, rd);
                }
                if (fin) {
                    if (ANY) dt_store_shadow(sq, ray, T.best, accum); else dt_store_closest(q, ray, T.best);
                    ray = -1;
                }
            }
#else
            if (ray >= 0) {
#ifdef DT_TIMELINE
                tl_steps++;
#endif
                if (dt_trav_step<ANY, WW>(T, stack, S, (ANY ? sq.o_time : q.o_time) + ray, (ANY ? sq.d_tmax : q.d_tmax) + ray)) {
#ifdef DT_TIMELINE
                    atomicAdd(&g_dt_steps_hist[ANY ? 1 : 0][min(63, tl_steps / 8)], 1u); tl_steps = 0;
#endif
                    if (ANY) dt_store_shadow(sq, ray, T.best, accum); else dt_store_closest(q, ray, T.best);
                    ray = -1;
                }
            }
#endif
            act = __ballot_sync(FULL, ray >= 0);
            if (act == 0u) break;
            if (!drained && __popc(act) < refill_threshold) break;
        }
    }
#ifdef DT_TIMELINE
    if (lane == 0) {
        const unsigned int k = atomicAdd(&g_dt_tl_count, 1u);
        if (k < DT_TL_MAX) { g_dt_tl[k * 4] = ((unsigned long long)(ANY ? 1 : 0) << 32) | (unsigned int)n; g_dt_tl[k * 4 + 1] = tl_t0; g_dt_tl[k * 4 + 2] = tl_td; g_dt_tl[k * 4 + 3] = dt_now(); }
    }
#endif
}

// occlusion query with an explicit result array (dt_trace_occluded)
__global__ void __launch_bounds__(128) k_traverse_occluded(DtSceneDev S, const float4* o_time, const float4* d_tmax, int n, uint8_t* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 o = o_time[i], d = d_tmax[i];
    DtHit h;
    dt_trace<true>(S, V(o.x, o.y, o.z), V(d.x, d.y, d.z), o.w, d.w, h);
    out[i] = h.shape >= 0 ? 1 : 0;
}

// ------------------------------------------------------------------ shade
struct DtChild {
    v3 o, d; float mb;
    v3 W; float n_medium;
    v3 thr; float beer_thr;
    int depth, beer_mat; uint32_t rng_key; int flags;
    v3 miss;
};

// slot < 0: allocate one here; otherwise a slot obtained earlier (dt_agg_inc / dt_agg_reserve on counters.next), so that the
// atomic's round trip overlaps the computation of the child
__device__ __forceinline__ int dt_emit_child(const DtRayQueue& out, float4* out_miss, const DtShadeCounters& counters, int capacity, uint32_t pix, const DtChild& c, int slot = -1) {
    if (slot < 0) slot = dt_agg_inc(counters.next);
    if (slot >= capacity) { atomicAdd(counters.overflow, 1); return -1; }
    out.o_time[slot] = make_float4(c.o.x, c.o.y, c.o.z, c.mb);
    out.d_tmax[slot] = make_float4(c.d.x, c.d.y, c.d.z, CUDART_INF_F);
    out.pixel[slot] = pix;
    out.weight_n[slot] = make_float4(c.W.x, c.W.y, c.W.z, c.n_medium);
    out.thr_beer[slot] = make_float4(c.thr.x, c.thr.y, c.thr.z, c.beer_thr);
    out.misc[slot] = make_int4(c.depth, c.beer_mat, (int)c.rng_key, c.flags);
    if (out_miss) out_miss[slot] = make_float4(c.miss.x, c.miss.y, c.miss.z, 0.f);
    return slot;
}

// slot < 0: allocate one here; otherwise a slot reserved with dt_agg_reserve
__device__ __forceinline__ void dt_emit_shadow(const DtShadowQueue& sq, const DtShadeCounters& counters, int capacity, v3 o, v3 d, float mb, float tmax,
                                               v3 contrib, uint32_t pix, int defer_slot, int defer_light, int slot = -1) {
    if (slot < 0) slot = dt_agg_inc(counters.shadow);
    if (slot >= capacity) { atomicAdd(counters.overflow, 1); return; }
    sq.d_tmax[slot] = make_float4(d.x, d.y, d.z, tmax);
    if (tmax < 0.0f) {                                       // an entry that will not be traced (dt_shade_ray): the marker alone
        if (sq.defer) sq.defer[slot] = make_int2(-1, -1);
        return;
    }
    sq.o_time[slot] = make_float4(o.x, o.y, o.z, mb);
    sq.contrib_pix[slot] = make_float4(contrib.x, contrib.y, contrib.z, __int_as_float((int)pix));
    if (sq.defer) sq.defer[slot] = make_int2(defer_slot, defer_light);
}

// A sphere with a normal map keeps the normal hitInfo held BEFORE Sphere::Intersect accepted the hit (sphere.cpp:95-115: the branch
// only reads the texture, its body is commented out): the normal of the closest hit among the shapes scanned earlier (each
// acceptance overwrites hitInfo.normal and lowers minT, so the last writer is the nearest of them), or (0,0,0) of a fresh HitInfo.
// If that shape is such a sphere too, it passed on ITS predecessors' normal through its own transform.  Rare and slow by design
// (one extra prefix traversal per such hit); not inlined so that the shading kernel's register allocation does not see it.
__device__ __noinline__ void dt_stale_normal(const DtSceneDev& S, float ox, float oy, float oz, float dx, float dy, float dz, float mb, int shape_limit, int smooth, float* out) {
    const v3 o = V(ox, oy, oz), d = V(dx, dy, dz);
    int chain[4], n_chain = 0;
    v3 normal = V(0.f, 0.f, 0.f);
    for (;;) {
        DtHit h;
        dt_trace_prefix(S, o, d, mb, shape_limit, h);
        if (h.shape < 0) break;
        const DtShapeDev& sh = S.shapes[h.shape];
        if (sh.kind == DT_SHAPE_SPHERE && sh.tex_normal >= 0) {
            if (n_chain == 4) break;
            chain[n_chain++] = h.shape; shape_limit = h.shape;
            continue;
        }
        v3 lo = apply_transform(sh.inv, o, 1.0f);
        v3 ld = apply_transform(sh.inv, d, 0.0f);
        if (sh.has_motion_blur) lo = vadd(lo, vscale(F3(sh.motion_blur), mb));
        DtSurface sf;
        if (sh.kind == DT_SHAPE_SPHERE) sphere_surface(S, sh, h.t, lo, ld, V(0.f, 0.f, 0.f), sf);
        else mesh_surface(S, sh, h.face, h.t, h.beta, h.gamma, lo, ld, smooth != 0, sf);
        normal = sf.normal;
        break;
    }
    while (n_chain > 0) normal = vunit(apply_transform(S.shapes[chain[--n_chain]].invT, normal, 0.0f));
    out[0] = normal.x; out[1] = normal.y; out[2] = normal.z;
}


// One thread per traced ray: Raytracer::PerPixel miss handling (raytracer.cpp:49-62) and PerformShading
// (:65-134) with ComputeGlobalIllumination (:135-191), SampleDirectLighting (:701-806), mirror / conductor /
// dielectric children (:208-472) turned into queue emissions.
template <bool STALE>
__device__ __forceinline__ void dt_shade_ray(const int i, const DtSceneDev& S, const DtCamDev& cam, const DtRayQueue& in, const float4* in_miss,
                                             const DtRayQueue& out, float4* out_miss, int out_capacity,
                                             const DtShadowQueue& sq, int shadow_capacity, const DtShadeCounters& counters, float4* accum, int* block_dead) {
    const uint32_t pix = in.pixel[i];
    if (pix == DT_DEAD_PIXEL) return;
    const float4 o4 = in.o_time[i], d4 = in.d_tmax[i];
    const float4 h0 = in.hit0[i];
    const float4 wn = in.weight_n[i];
    const int4 misc = in.misc[i];
    const v3 o = V(o4.x, o4.y, o4.z), d = V(d4.x, d4.y, d4.z);
    const float mb = o4.w;
    v3 W = V(wn.x, wn.y, wn.z);
    const float n_medium = wn.w;
    const int hit_shape = __float_as_int(h0.w);
    const int flags = misc.w;
    const bool primary = (flags & DT_FLAG_PRIMARY) != 0;
    const int depth = primary ? S.max_recursion_depth : misc.x;

    if (hit_shape < 0) {
        if (primary) {
            v3 c;
            if (S.bg_texture >= 0) {
                const int x = (int)(pix % (uint32_t)cam.width) + (misc.y & 1), y = (int)(pix / (uint32_t)cam.width) + ((misc.y >> 1) & 1);   // coordX / coordY of RenderPixel
                c = tex_rgb_sample(S, S.textures[S.bg_texture], x / (float)cam.width, y / (float)cam.height);
            } else if (S.n_env_lights > 0) c = env_sample(S, 0, d);
            else c = V((float)S.background_color[0], (float)S.background_color[1], (float)S.background_color[2]);
            dt_accum(accum, pix, vmul(W, c));
        } else if ((flags & DT_FLAG_ENV_ON_MISS) && in_miss) {
            const float4 m = in_miss[i];
            dt_accum(accum, pix, vmul(W, V(m.x, m.y, m.z)));
        }
        return;
    }

    const float t = h0.x;
    const DtShapeDev& sh = S.shapes[hit_shape];
    const v3 hitPoint = vadd(o, vscale(d, t));                                   // raytracer.cpp:69
    // ---- surface at the winning hit ----
    DtSurface sf;
    {
        v3 lo = apply_transform(sh.inv, o, 1.0f);
        v3 ld = apply_transform(sh.inv, d, 0.0f);
        if (sh.has_motion_blur) lo = vadd(lo, vscale(F3(sh.motion_blur), mb));
        if (sh.kind == DT_SHAPE_SPHERE) {
            float stale[3] = {0.f, 0.f, 0.f};
            if (STALE && sh.tex_normal >= 0) dt_stale_normal(S, o.x, o.y, o.z, d.x, d.y, d.z, mb, hit_shape, cam.smooth_shading, stale);
            sphere_surface(S, sh, t, lo, ld, V(stale[0], stale[1], stale[2]), sf);
        } else mesh_surface(S, sh, in.hit_face[i], t, h0.y, h0.z, lo, ld, cam.smooth_shading != 0, sf);
    }
    const v3 normal = sf.normal;
    const dt_material mat = S.materials[sh.material - 1];

    // Beer's law applied by the parent to everything this child returns (raytracer.cpp:306-309,345-349,398-402)
    const float4 tb = in.thr_beer[i];
    if (tb.w > 0.0f && n_medium > tb.w) {
        const dt_material& pm = S.materials[misc.y - 1];
        W = V(W.x * expf(-pm.absorption_coefficient[0] * t), W.y * expf(-pm.absorption_coefficient[1] * t), W.z * expf(-pm.absorption_coefficient[2] * t));
    }

    // A path whose weight has underflowed to exactly zero adds exact zeros from here on (every term below is W times something).
    // Under Russian roulette, whose pure-GI chains never end (raytracer.cpp:137-147: the throughput stays 1), such paths are what
    // the last thousands of waves of a frame consist of: rays caught inside the closed 10 M-triangle mesh keep bouncing, with W = 0
    // after ~100 bounces.  They end here.  Bounded-depth frames keep them (their ray counts stay the reference's); the one
    // observable difference is that a NaN produced after the underflow no longer poisons the pixel.
    if (cam.path_tracing && cam.russian_roulette && !cam.keep_weightless && W.x == 0.0f && W.y == 0.0f && W.z == 0.0f) return;

    const v3 eye = primary ? F3(cam.position) : o;
    const v3 w_o = vunit(vsub(eye, hitPoint));
    const float vac = 1.00001f;
    const bool inside = n_medium > vac;

    if (mat.type == DT_MAT_EMISSIVE) {                                            // raytracer.cpp:81-84
        dt_accum(accum, pix, vmul(W, vscale(vscale(F3(mat.radiance), 2.0f), DT_PI_F)));
        return;
    }
    if (sh.tex_replace_all >= 0) {                                                // raytracer.cpp:87-89
        dt_accum(accum, pix, vmul(W, tex_rgb_sample(S, S.textures[sh.tex_replace_all], sf.u, sf.v)));
        return;
    }

    DtRng rng; rng.key = (uint32_t)misc.z; rng.ctr = 0;
    v3 thr = V(tb.x, tb.y, tb.z);
    int gi_slot = -1;
    bool gi_weightless = false, live_deferred = false;     // the GI child's weight is exactly zero; a traced mesh-light sample waits for its hit

    // Queue slots are requested up front so that the atomics' round trips overlap the sampling / BRDF arithmetic below.
    // Every point / area / directional / spot / mesh light emits exactly one shadow ray per shaded hit: one reservation.
    const bool sampleDirect = !cam.path_tracing || cam.nee;
    const bool direct = !inside && sampleDirect;
    const int n_shadow = direct ? S.n_point_lights + S.n_area_lights + S.n_directional_lights + S.n_spot_lights + S.n_mesh_lights : 0;
    int shadow_stride = 0;
    int shadow_slot = n_shadow > 0 ? dt_agg_reserve(counters.shadow, n_shadow, shadow_stride) : 0;

    // kd / ks of this hit (Get{Diffuse,Specular}ReflectanceCoeff, raytracer.cpp:478-539): light-independent, evaluated once
    const v3 kd = reflectance_coeff(S, sh, mat, hitPoint, sf.u, sf.v, false);
    const v3 ks = reflectance_coeff(S, sh, mat, hitPoint, sf.u, sf.v, true);

    // ---- ComputeGlobalIllumination (raytracer.cpp:135-191), then ambient + SampleDirectLighting (:98-108, 701-806) ----
    // ONE loop over "incoming directions": item -1 is the GI sample, items 0.. are the lights in the reference's order (point, area,
    // environment, directional, spot, mesh), so the RNG draws and the throughput updates happen in the reference's sequence while
    // Shade() -- the BRDF switch, by far the largest piece of code in this kernel -- exists once instead of seven times
    // (the kernel was 38 000 SASS instructions and a fifth of its stall samples were instruction-cache misses).
    const int e_point = S.n_point_lights, e_area = e_point + S.n_area_lights, e_env = e_area + S.n_env_lights,
              e_dir = e_env + S.n_directional_lights, e_spot = e_dir + S.n_spot_lights, e_mesh = e_spot + S.n_mesh_lights;
    v3 local = vmul(F3(S.ambient_light), F3(mat.ambient));
    const v3 so = vadd(hitPoint, vscale(normal, S.shadow_ray_epsilon));
    for (int item = cam.path_tracing ? -1 : 0; item < (direct ? e_mesh : 0); item++) {
        v3 w_i, E;
        float lightT = CUDART_INF_F;
        int kind;                       // 0 GI child, 1 shadow ray, 2 added unshadowed (environment light)
        int defer_light = -1;
        v3 child_d = V(0.f, 0.f, 0.f);
        if (item < 0) {
            bool go = true;
            if (cam.russian_roulette) {
                float probTest = rng01(rng);
                float maxT = fmaxf(thr.x, fmaxf(thr.x, thr.z));
                if (probTest > maxT && depth <= 0) go = false;
                else if (depth < -DT_RR_MAX_BOUNCES) go = false;           // see DT_RR_MAX_BOUNCES
                else thr = vdiv(thr, maxT);
            } else if (depth <= 0) go = false;
            if (!go) continue;
            gi_slot = dt_agg_inc(counters.next);
            float rand1 = rng01(rng), rand2 = rng01(rng);
            float phi = (float)(2 * DT_PI * rand1);
            float theta = cam.importance_sampling ? asinf(sqrtf(rand2)) : acosf(rand2);
            v3 u, v;
            orthonormal_basis(normal, u, v);
            v3 nd = vadd(vadd(vscale(vscale(u, sinf(theta)), cosf(phi)), vscale(normal, cosf(theta))), vscale(vscale(v, sinf(theta)), sinf(phi)));
            nd = vunit(nd);
            child_d = nd; w_i = nd; E = V(1.f, 1.f, 1.f); kind = 0;
        } else if (item < e_point) {
            const dt_point_light& L = S.point_lights[item];
            v3 lp = F3(L.position);
            v3 dir = vsub(lp, hitPoint);
            lightT = vlen(dir);
            w_i = vdiv(dir, lightT);                                              // makeUnit(lp - hitPoint) == the same three divisions
            E = vdiv(F3(L.intensity), (lightT * lightT));
            kind = 1;
        } else if (item < e_area) {
            const dt_area_light& L = S.area_lights[item - e_point];
            float offU = -0.5f + rng01(rng), offV = -0.5f + rng01(rng);          // areaLight.h:34-40
            v3 sp = vadd(vadd(F3(L.position), vscale(F3(L.u), (L.extent * offU))), vscale(F3(L.v), (L.extent * offV)));
            v3 dir = vsub(sp, hitPoint);
            lightT = vlen(dir);
            w_i = vdiv(dir, lightT);
            float dSqr = lightT * lightT;
            float lCos = vdot(F3(L.normal), vneg(w_i));
            if (lCos < 0) lCos = vdot(F3(L.normal), w_i);
            float area = L.extent * L.extent;
            E = vscale(F3(L.radiance), (area * lCos / dSqr));
            kind = 1;
        } else if (item < e_env) {                                                // no shadow ray (raytracer.cpp:741-755)
            // SphericalEnvironmentLight::GetDirection (sphericalEnvironmentLight.h:36-65) rejection-samples the cube [-1,1]^3 until the
            // candidate lies in the unit ball on the normal's side and returns it UN-normalised: a uniform point of the half ball.  A
            // per-lane rejection loop runs as long as the unluckiest lane of the warp (12 rounds at 26 % acceptance, 9 live lanes:
            // a quarter of k_shade's issued instructions on config 5), so the same distribution is drawn directly: radius u^(1/3),
            // uniform direction, mirrored into the normal's half space.
            const v3 nn = vunit(normal);
            const float ur = rng01(rng), uz = rng01(rng), up = rng01(rng);
            const float rad = cbrtf(ur), cz = 1.0f - 2.0f * uz, sz = sqrtf(fmaxf(0.0f, 1.0f - cz * cz));
            float sp, cp;
            sincosf((float)(2 * DT_PI) * up, &sp, &cp);
            v3 cand = V(rad * sz * cp, rad * sz * sp, rad * cz);
            if (vdot(nn, cand) < 0.0f) cand = V(-cand.x, -cand.y, -cand.z);
            E = env_sample(S, item - e_area, cand);
            w_i = normal;
            kind = 2;
        } else if (item < e_dir) {
            const dt_directional_light& L = S.directional_lights[item - e_env];
            w_i = vneg(F3(L.dir));
            E = F3(L.radiance);
            kind = 1;
        } else if (item < e_spot) {
            const dt_spot_light& L = S.spot_lights[item - e_dir];
            v3 dir = vsub(F3(L.pos), hitPoint);
            lightT = vlen(dir);
            w_i = vdiv(dir, lightT);
            E = spot_irradiance(L, hitPoint);
            kind = 1;
        } else {
            const dt_mesh_light& L = S.mesh_lights[item - e_spot];
            const DtShapeDev& lsh = S.shapes[L.shape];
            const DtMeshDev& lm = S.meshes[lsh.mesh];
            int fi = (int)(rng01(rng) * lm.n_faces);                              // meshLight.h:27-47 (+P2)
            if (fi >= lm.n_faces) fi = lm.n_faces - 1;
            const DtFaceDev fc = S.faces[lm.face_base + fi];
            float rand1 = rng01(rng), rand2 = rng01(rng);
            const float* vp = S.verts + (size_t)lm.vert_base * 3;
            v3 a = F3(vp + (size_t)fc.v0 * 3), b = F3(vp + (size_t)fc.v1 * 3), cc = F3(vp + (size_t)fc.v2 * 3);
            v3 qq = vadd(vscale(b, (1 - rand2)), vscale(cc, rand2));
            float sr = sqrtf(rand1);
            v3 pos = vadd(vscale(a, (1 - sr)), vscale(qq, sr));
            pos = apply_transform(lsh.fwd, pos, 1.0f);
            v3 dir = vsub(pos, hitPoint);
            lightT = vlen(dir);
            w_i = vdiv(dir, lightT);
            E = vscale(vscale(vscale(F3(L.radiance), fc.light_weight), 2.f), DT_PI_F);
            // hitMeshLightId (raytracer.cpp:91-95,107,781): this light is skipped when the GI child of this very
            // hit lands on the emissive shape with the same id -> decided one wave later (deferred entry).
            defer_light = L.id;
            kind = 1;
        }
        v3 res = V(1.f, 1.f, 1.f);
        const v3 c = shade_term(mat, S, kd, ks, normal, w_i, w_o, E, &res);
        if (kind == 0) {
            DtChild ch;
            ch.o = vadd(hitPoint, vscale(normal, (float)0.0001)); ch.d = child_d; ch.mb = mb;
            ch.W = vmul(W, vscale(vscale(c, 2.0f), DT_PI_F));
            ch.n_medium = n_medium; ch.thr = thr; ch.beer_thr = 0.f;
            ch.depth = depth - 1; ch.beer_mat = 0; ch.rng_key = dt_hash(rng.key, 0xA511E9B3u); ch.flags = 0; ch.miss = V(0, 0, 0);
            gi_slot = dt_emit_child(out, out_miss, counters, out_capacity, pix, ch, gi_slot);
            gi_weightless = ch.W.x == 0.0f && ch.W.y == 0.0f && ch.W.z == 0.0f;
        }
        if (mat.brdf >= 0) thr = vmul(thr, res);                                  // Shade(): ray.throughput *= res
        if (kind == 1) {
            // A shadow ray whose contribution is EXACTLY zero -- the light is below the surface's horizon, the lobe or the irradiance
            // vanishes -- cannot change the pixel whatever it hits.  Path-traced frames do not trace it: its (already reserved) queue
            // slot gets tmax = -1, which the any-hit kernels skip.  Whitted frames keep it (ray counts stay the reference's), and so
            // does DT_FLAG_KEEP_WEIGHTLESS_PATHS.
            const v3 contrib = vmul(W, c);
            bool dead = false;
            if (cam.path_tracing && !cam.keep_weightless) {      // (frame-uniform: Whitted frames do not pay for the vote)
                dead = contrib.x == 0.0f && contrib.y == 0.0f && contrib.z == 0.0f;
                // counted per block in shared memory (a per-thread counter carried through this function costs k_shade 7 %)
                const unsigned dm = __ballot_sync(__activemask(), dead);
                if (dead && (threadIdx.x & 31) == __ffs(dm) - 1) atomicAdd(block_dead, __popc(dm));
            }
            if (!dead && defer_light >= 0) live_deferred = true;
            dt_emit_shadow(sq, counters, shadow_capacity, so, w_i, mb, dead ? -1.0f : lightT, contrib, pix, defer_light >= 0 ? gi_slot : -1, defer_light, shadow_slot);
            shadow_slot += shadow_stride;
        } else if (kind == 2) local = vadd(local, c);
    }
    if (direct) dt_accum(accum, pix, vmul(W, local));
    // A GI child of exactly zero weight (a black surface: the mirror sphere of configs 4 / 5) would be traced once and dropped when it
    // is shaded; unless a traced mesh-light sample of this hit needs to know what it hits (deferred NEE), it is not traced at all.
    if (gi_weightless && !live_deferred && gi_slot >= 0 && cam.russian_roulette && !cam.keep_weightless) {
        out.pixel[gi_slot] = DT_DEAD_PIXEL;
        atomicAdd(block_dead + 1, 1);              // shared memory, rare
    }

    if (depth <= 0) return;        // all three recursive helpers start with `if(recDepth <= 0) return 0`

    if (mat.type == DT_MAT_MIRROR) {                                              // raytracer.cpp:442-472
        const int slot = dt_agg_inc(counters.next);
        DtChild c;
        c.d = reflect_dir(rng, normal, w_o, mat.roughness);
        c.o = vadd(hitPoint, vscale(normal, S.shadow_ray_epsilon)); c.mb = mb;
        c.W = vmul(W, F3(mat.mirror)); c.n_medium = 1.0f; c.thr = thr; c.beer_thr = 0.f;
        c.depth = depth - 1; c.beer_mat = 0; c.rng_key = dt_hash(rng.key, 0x1B873593u); c.flags = 0; c.miss = V(0, 0, 0);
        if (S.n_env_lights > 0) { c.flags |= DT_FLAG_ENV_ON_MISS; c.miss = env_sample(S, 0, c.d); }
        dt_emit_child(out, out_miss, counters, out_capacity, pix, c, slot);
    } else if (mat.type == DT_MAT_CONDUCTOR) {                                    // raytracer.cpp:208-254
        v3 dd = vneg(w_o);
        float cosTheta = -vdot(dd, normal);
        float n2 = mat.refractive_index, k2 = mat.conductor_absorption_index;
        float n2k2 = n2 * n2 + k2 * k2;
        float n2cosTheta2 = 2 * n2 * cosTheta;
        float cosThetaSqr = cosTheta * cosTheta;
        float rs = (n2k2 - n2cosTheta2 + cosThetaSqr) / (n2k2 + n2cosTheta2 + cosThetaSqr);
        float rp = (n2k2 * cosThetaSqr - n2cosTheta2 + 1) / (n2k2 * cosThetaSqr + n2cosTheta2 + 1);
        float reflectRatio = (float)(0.5 * (double)(rs + rp));
        if (reflectRatio > 0.0001) {
            DtChild c;
            c.d = reflect_dir(rng, normal, w_o, mat.roughness);
            c.o = vadd(hitPoint, vscale(normal, S.shadow_ray_epsilon)); c.mb = mb;
            c.W = vscale(vmul(W, F3(mat.mirror)), reflectRatio); c.n_medium = 1.0f; c.thr = thr; c.beer_thr = 0.f;
            c.depth = depth - 1; c.beer_mat = 0; c.rng_key = dt_hash(rng.key, 0x2C1B3C6Du); c.flags = 0; c.miss = V(0, 0, 0);
            dt_emit_child(out, out_miss, counters, out_capacity, pix, c);
        }
    } else if (mat.type == DT_MAT_DIELECTRIC) {                                   // raytracer.cpp:261-415
        float n1 = n_medium, n2 = mat.refractive_index;
        v3 dd = vneg(w_o);
        v3 mn = normal;
        float cosTheta = -vdot(dd, mn);
        const bool isEntering = cosTheta > 0.f;
        float objN = n2;
        if (!isEntering) { n1 = n2; n2 = 1.0f; objN = 1.0f; cosTheta = fabsf(cosTheta); mn = vneg(mn); }
        float r = n1 / n2;
        float sinThetaSqr = 1 - (cosTheta * cosTheta);
        float criticalTerm = r * r * sinThetaSqr;
        const int mat_id = sh.material;
        if (criticalTerm > 1) {
            DtChild c;
            c.d = reflect_dir(rng, mn, w_o, mat.roughness);
            c.o = vadd(hitPoint, vscale(mn, S.shadow_ray_epsilon)); c.mb = mb;
            c.W = W; c.n_medium = n_medium; c.thr = thr; c.beer_thr = 1.0001f;
            c.depth = depth - 1; c.beer_mat = mat_id; c.rng_key = dt_hash(rng.key, 0x3D4D51CBu); c.flags = 0; c.miss = V(0, 0, 0);
            dt_emit_child(out, out_miss, counters, out_capacity, pix, c);
        } else {
            int stride2 = 0;
            const int slot2 = dt_agg_reserve(counters.next, 2, stride2);          // reflected + refracted child
            float cosPhi = sqrtf(1 - criticalTerm);
            float n2cosTheta = n2 * cosTheta;
            float n1cosPhi = n1 * cosPhi;
            float rparallel = (n2cosTheta - n1cosPhi) / (n2cosTheta + n1cosPhi);
            float rperp = (n1 * cosTheta - n2 * cosPhi) / (n1 * cosTheta + n2 * cosPhi);
            float rReflect = (rparallel * rparallel + rperp * rperp) / 2;
            float rRefract = 1 - rReflect;
            const float child_n = isEntering ? objN : 1.0f;
            DtChild c;
            c.d = reflect_dir(rng, mn, w_o, mat.roughness);
            const v3 refl_dir = c.d;
            c.o = vadd(hitPoint, vscale(mn, S.shadow_ray_epsilon)); c.mb = mb;
            c.W = vscale(W, rReflect); c.n_medium = child_n; c.thr = thr; c.beer_thr = 1.00001f;
            c.depth = depth - 1; c.beer_mat = mat_id; c.rng_key = dt_hash(rng.key, 0x4CF5AD43u); c.flags = 0; c.miss = V(0, 0, 0);
            v3 env = V(0, 0, 0);
            if (S.n_env_lights > 0) { env = env_sample(S, 0, refl_dir); c.flags |= DT_FLAG_ENV_ON_MISS; c.miss = env; }
            dt_emit_child(out, out_miss, counters, out_capacity, pix, c, slot2);

            v3 w_t = vsub(vscale(vadd(dd, vscale(mn, cosTheta)), r), vscale(mn, cosPhi));
            if (mat.roughness > 0.001) {
                v3 u, v;
                orthonormal_basis(w_t, u, v);
                float psi1 = rng01(rng) - 0.5f, psi2 = rng01(rng) - 0.5f;
                w_t = vunit(vadd(w_t, vscale(vadd(vscale(u, psi1), vscale(v, psi2)), mat.roughness)));
            } else w_t = vunit(w_t);
            c.d = w_t;
            c.o = vadd(hitPoint, vscale(vneg(mn), S.shadow_ray_epsilon));
            c.W = vscale(W, rRefract); c.beer_thr = 1.001f;
            c.rng_key = dt_hash(rng.key, 0x5A27B1E9u);
            // a refracted ray that misses reads the environment along the REFLECTED direction (raytracer.cpp:408)
            dt_emit_child(out, out_miss, counters, out_capacity, pix, c, slot2 + stride2);
        }
    }
}

// Grid-stride launch: the wave size lives in device memory (n_ptr) so consecutive waves need no host round trip.
// `perm` (optional) is the material-sorted order of the wave produced by the sort stage below.
#ifndef DT_SHADE_MINBLOCKS
#define DT_SHADE_MINBLOCKS 4       // 128 registers: measured best on config 4 (shade 162 ms at 193 regs, 133 at 170, 124 at 128, 125 at 102)
#endif
// STALE: the scene has a sphere with a normal map (dt_stale_normal); a separate instantiation so that every other scene runs the kernel without that call
template <bool STALE>
__global__ void __launch_bounds__(128, DT_SHADE_MINBLOCKS) k_shade(DtSceneDev S, DtCamDev cam, DtRayQueue in, const float4* in_miss, const int* n_ptr, int n_fixed, const int* perm,
                                               DtRayQueue out, float4* out_miss, int out_capacity,
                                               DtShadowQueue sq, int shadow_capacity, DtShadeCounters counters, float4* accum) {
    const int n = n_ptr ? *n_ptr : n_fixed;
    // The lanes of a warp take the SAME number of trips and meet again after every hit: a lane whose hit returns early (a miss, an
    // emissive surface, a path whose weight is zero) would otherwise run ahead into its next hits on its own -- nothing reconverges
    // a warp at a loop back-edge -- and the warp degenerates into 32 single-lane executions of this kernel (measured on config 5
    // once a fifth of the hits returned early: k_shade 2.3x slower in the device-resident loop than with one hit per thread).
    const int lane = threadIdx.x & 31;
    __shared__ int s_dead[2];                    // shadow-queue / closest-hit queue entries this block marked "not traced"
    if (threadIdx.x < 2) s_dead[threadIdx.x] = 0;
    __syncthreads();
    for (int base = blockIdx.x * blockDim.x + threadIdx.x - lane; base < n; base += gridDim.x * blockDim.x) {
        const int j = base + lane;
        if (j < n) dt_shade_ray<STALE>(perm ? perm[j] : j, S, cam, in, in_miss, out, out_miss, out_capacity, sq, shadow_capacity, counters, accum, s_dead);
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_dead[0] > 0) atomicAdd(counters.shadow_dead, (unsigned long long)s_dead[0]);
    if (threadIdx.x == 1 && s_dead[1] > 0) atomicAdd(counters.closest_dead, (unsigned long long)s_dead[1]);
}

// ------------------------------------------------------------------ sort / compact by material
// Between closest-hit and shade: a counting sort of the wave by shading key (0 = dead slot or miss, 1 + material id of
// the hit shape otherwise), so that the lanes of a shading warp run the same material / BRDF / texture branches
// (PerformShading's switch over Material::type and BRDF, raytracer.cpp:65-134).  Two small launches: histogram,
// stable-per-block scatter of ray indices into `perm` (each block scans the <= 256 bin counts itself).
#define DT_SORT_BINS 256
__device__ __forceinline__ int dt_sort_key(const DtSceneDev& S, const DtRayQueue& q, int i) {
    if (q.pixel[i] == DT_DEAD_PIXEL) return 0;
    const int shape = __float_as_int(q.hit0[i].w);
    if (shape < 0) return 0;
    const int m = S.shapes[shape].material;
    return 1 + (m < 0 ? 0 : m % (DT_SORT_BINS - 1));
}
__global__ void __launch_bounds__(256) k_sort_hist(DtSceneDev S, DtRayQueue q, const int* n_ptr, int n_fixed, int* hist) {
    __shared__ int h[DT_SORT_BINS];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int n = n_ptr ? *n_ptr : n_fixed;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int k = dt_sort_key(S, q, i);
        q.sort_key[i] = (uint32_t)k;
        // a wave holds a handful of distinct keys: the lanes that share one send ONE shared-memory atomic, not up to 32 serialised ones
        const unsigned peers = __match_any_sync(__activemask(), k);
        if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&h[k], __popc(peers));
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}
// Each block turns the bin counts into exclusive offsets itself (256 values: cheaper than a third launch between the
// histogram and the scatter), takes contiguous chunks of the wave, reserves one range per bin through the running cursors
// (zeroed with the histogram) and writes its rays in chunk order.
__global__ void __launch_bounds__(256) k_sort_scatter(DtRayQueue q, const int* n_ptr, int n_fixed, const int* hist, int* cursors, int* perm) {
    __shared__ int cnt[DT_SORT_BINS], base[DT_SORT_BINS], offs[DT_SORT_BINS];
    static_assert(DT_SORT_BINS == 256, "one thread per bin");
    const int n = n_ptr ? *n_ptr : n_fixed;
    const int chunk = 256 * 8;
    if (blockIdx.x * chunk >= n) return;
    {
        const int t = threadIdx.x, v = hist[t];
        offs[t] = v;
        __syncthreads();
        for (int d = 1; d < DT_SORT_BINS; d <<= 1) {
            const int x = t >= d ? offs[t - d] : 0;
            __syncthreads();
            offs[t] += x;
            __syncthreads();
        }
        offs[t] -= v;                                                              // exclusive
    }
    for (int c0 = blockIdx.x * chunk; c0 < n; c0 += gridDim.x * chunk) {
        cnt[threadIdx.x] = 0;
        __syncthreads();
        int key[8], rank[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = c0 + k * 256 + threadIdx.x;
            key[k] = i < n ? (int)q.sort_key[i] : -1;
            // one atomic per distinct key of the warp (see k_sort_hist); the lanes that share a key rank themselves by lane index
            const unsigned peers = __match_any_sync(0xFFFFFFFFu, key[k]);
            const int leader = __ffs(peers) - 1;
            int base_rank = 0;
            if ((int)(threadIdx.x & 31) == leader && key[k] >= 0) base_rank = atomicAdd(&cnt[key[k]], __popc(peers));
            base_rank = __shfl_sync(0xFFFFFFFFu, base_rank, leader);
            rank[k] = base_rank + __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
        }
        __syncthreads();
        if (cnt[threadIdx.x]) base[threadIdx.x] = offs[threadIdx.x] + atomicAdd(&cursors[threadIdx.x], cnt[threadIdx.x]);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; k++) if (key[k] >= 0) perm[base[key[k]] + rank[k]] = c0 + k * 256 + threadIdx.x;
        __syncthreads();
    }
}

// ------------------------------------------------------------------ sort by hit cell (+ material)
// Path-traced waves are incoherent: after one diffuse bounce the queue order is unrelated to where the rays are, the lanes of a
// traversal warp walk unrelated subtrees (11-14 live threads per issued instruction on config 5) and every node is a fresh L2 /
// HBM sector.  The children and the shadow rays of a hit START at the hit point, and k_shade emits them in the order it
// processes the hits, so ordering the HITS in space before shading hands the next closest-hit pass and this wave's shadow pass
// warps of rays with neighbouring origins (shadow rays: neighbouring origins AND the same light), and groups the shading warps'
// surface / material reads as a side effect.  Key, most significant first:
//     Morton code of the hit cell (b bits per axis, b grows with the wave: 1..DT_SSORT_MAX_AXIS_BITS) | material class (4 bits)
//     | 9 more Morton bits (3 per axis)
// The upper part (<= 22 bits) is counting-sorted through global atomics (histogram, scan, scatter: the order inside a bin is
// arbitrary); k_ssort_refine then sorts chunks of 2048 consecutive entries by the full key in shared memory, which orders the
// ~100-300 hits of a coarse cell by the finer cell.  Misses and dead slots (class 0) are spread over the bins of their class
// so that they do not serialise on one counter.
#define DT_SSORT_CLASS_BITS 4
#define DT_SSORT_MAX_AXIS_BITS 6                  // 16 classes << 18 = 4 Mi bins: a 16 MiB counter table, L2-resident
#define DT_SSORT_LO_BITS 9
#define DT_SSORT_SCAN_BLOCK 4096                  // bins per block of the scan kernels
#define DT_SSORT_MAX_BINS (1 << (DT_SSORT_CLASS_BITS + 3 * DT_SSORT_MAX_AXIS_BITS))
#define DT_SSORT_CHUNK 2048
struct DtSpatialSort {
    int* hist;                    // DT_SSORT_MAX_BINS counters, all zero between waves (k_ssort_scan_b re-zeroes what a wave touched)
    int* cursor;                  // DT_SSORT_MAX_BINS running offsets
    int* blocksum;                // DT_SSORT_MAX_BINS / DT_SSORT_SCAN_BLOCK partial sums
    int max_axis_bits;            // <= DT_SSORT_MAX_AXIS_BITS (A/B knob)
    int refine;                   // run k_ssort_refine
};
// bits per axis of the coarse cell for a wave of n rays: ~1000 rays per 4^b, so that occupied cells (surfaces are 2-D) hold tens of rays
__device__ __forceinline__ int dt_ssort_axis_bits(int n, int max_bits) {
    int b = 1;
    while (b < max_bits && (1024 << (2 * (b + 1))) <= n) b++;
    return b;
}
__device__ __forceinline__ uint32_t dt_part1by2(uint32_t x) {      // spread the low 10 bits: bit k -> bit 3k
    x &= 0x3FFu;
    x = (x | (x << 16)) & 0x030000FFu;
    x = (x | (x << 8)) & 0x0300F00Fu;
    x = (x | (x << 4)) & 0x030C30C3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}
__device__ __forceinline__ uint32_t dt_ssort_key(const DtSceneDev& S, const DtRayQueue& q, int i, int b) {
    int shape = -1;
    float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q.pixel[i] != DT_DEAD_PIXEL) { h0 = q.hit0[i]; shape = __float_as_int(h0.w); }
    if (shape < 0) return (((uint32_t)i * 2654435761u) >> (32 - 3 * b)) << (DT_SSORT_CLASS_BITS + DT_SSORT_LO_BITS);          // class 0, any cell
    const int m = S.shapes[shape].material;
    const uint32_t cls = 1u + (uint32_t)(m < 1 ? 0 : (m - 1) % ((1 << DT_SSORT_CLASS_BITS) - 1));
    const float4 o = q.o_time[i], d = q.d_tmax[i];
    const int top = (1 << (b + 3)) - 1;
    const float fs = (float)(top + 1);
    const int qx = min(top, max(0, (int)((o.x + d.x * h0.x - S.sort_min[0]) * S.sort_scale[0] * fs)));
    const int qy = min(top, max(0, (int)((o.y + d.y * h0.x - S.sort_min[1]) * S.sort_scale[1] * fs)));
    const int qz = min(top, max(0, (int)((o.z + d.z * h0.x - S.sort_min[2]) * S.sort_scale[2] * fs)));
    const uint32_t hi = dt_part1by2((uint32_t)qx >> 3) | (dt_part1by2((uint32_t)qy >> 3) << 1) | (dt_part1by2((uint32_t)qz >> 3) << 2);
    const uint32_t lo = dt_part1by2((uint32_t)qx & 7u) | (dt_part1by2((uint32_t)qy & 7u) << 1) | (dt_part1by2((uint32_t)qz & 7u) << 2);
    return (((hi << DT_SSORT_CLASS_BITS) | cls) << DT_SSORT_LO_BITS) | lo;
}
__global__ void __launch_bounds__(256) k_ssort_hist(DtSceneDev S, DtRayQueue q, const int* n_ptr, int n_fixed, DtSpatialSort ss) {
    const int n = n_ptr ? *n_ptr : n_fixed;
    const int b = dt_ssort_axis_bits(n, ss.max_axis_bits);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t k = dt_ssort_key(S, q, i, b);
        q.sort_key[i] = k;
        atomicAdd(&ss.hist[k >> DT_SSORT_LO_BITS], 1);
    }
}
// exclusive scan of the bin counts, two launches: per-block sums, then every block adds up the sums of the blocks before it
__global__ void __launch_bounds__(256) k_ssort_scan_a(const int* n_ptr, int n_fixed, DtSpatialSort ss) {
    const int n = n_ptr ? *n_ptr : n_fixed;
    const int bins = 1 << (DT_SSORT_CLASS_BITS + 3 * dt_ssort_axis_bits(n, ss.max_axis_bits));
    const int base = blockIdx.x * DT_SSORT_SCAN_BLOCK;
    if (base >= bins) return;
    __shared__ int wsum[8];
    int v = 0;
    const int4* h4 = reinterpret_cast<const int4*>(ss.hist + base);
    if (base + DT_SSORT_SCAN_BLOCK <= bins) {
        for (int k = threadIdx.x; k < DT_SSORT_SCAN_BLOCK / 4; k += 256) { const int4 x = h4[k]; v += x.x + x.y + x.z + x.w; }
    } else {
        for (int k = threadIdx.x; base + k < bins; k += 256) v += ss.hist[base + k];
    }
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; w++) t += wsum[w]; ss.blocksum[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(256) k_ssort_scan_b(const int* n_ptr, int n_fixed, DtSpatialSort ss) {
    const int n = n_ptr ? *n_ptr : n_fixed;
    const int bins = 1 << (DT_SSORT_CLASS_BITS + 3 * dt_ssort_axis_bits(n, ss.max_axis_bits));
    const int base = blockIdx.x * DT_SSORT_SCAN_BLOCK;
    if (base >= bins) return;
    __shared__ int wsum[8];
    __shared__ int s_prefix;
    int v = 0;
    for (int k = threadIdx.x; k < (int)blockIdx.x; k += 256) v += ss.blocksum[k];
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; w++) t += wsum[w]; s_prefix = t; }
    __syncthreads();
    // thread t owns 16 consecutive bins
    const int per = DT_SSORT_SCAN_BLOCK / 256;
    int x[per], sum = 0;
    const int first = base + threadIdx.x * per;
#pragma unroll
    for (int k = 0; k < per; k++) { x[k] = first + k < bins ? ss.hist[first + k] : 0; sum += x[k]; }
    int incl = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += y; }
    __syncthreads();
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; w++) woff += wsum[w];
    int run = s_prefix + woff + incl - sum;
#pragma unroll
    for (int k = 0; k < per; k++) if (first + k < bins) { ss.cursor[first + k] = run; run += x[k]; ss.hist[first + k] = 0; }
}
__global__ void __launch_bounds__(256) k_ssort_scatter(DtRayQueue q, const int* n_ptr, int n_fixed, DtSpatialSort ss, int* perm) {
    const int n = n_ptr ? *n_ptr : n_fixed;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        perm[atomicAdd(&ss.cursor[q.sort_key[i] >> DT_SSORT_LO_BITS], 1)] = i;
}
// chunks of DT_SSORT_CHUNK consecutive entries of the coarse order, bitonic-sorted by the full key in shared memory
__global__ void __launch_bounds__(256) k_ssort_refine(DtRayQueue q, const int* n_ptr, int n_fixed, int* perm) {
    const int n = n_ptr ? *n_ptr : n_fixed;
    __shared__ unsigned long long e[DT_SSORT_CHUNK];
    for (int c0 = blockIdx.x * DT_SSORT_CHUNK; c0 < n; c0 += gridDim.x * DT_SSORT_CHUNK) {
        for (int k = threadIdx.x; k < DT_SSORT_CHUNK; k += 256) {
            const int j = c0 + k;
            unsigned long long v = ~0ull;
            if (j < n) { const int i = perm[j]; v = ((unsigned long long)q.sort_key[i] << 32) | (unsigned int)i; }
            e[k] = v;
        }
        __syncthreads();
        for (int size = 2; size <= DT_SSORT_CHUNK; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = threadIdx.x; t < DT_SSORT_CHUNK / 2; t += 256) {
                    const int lo = 2 * t - (t & (stride - 1));
                    const int hi = lo + stride;
                    const bool up = (lo & size) == 0;
                    const unsigned long long a = e[lo], b = e[hi];
                    if ((a > b) == up) { e[lo] = b; e[hi] = a; }
                }
                __syncthreads();
            }
        }
        for (int k = threadIdx.x; k < DT_SSORT_CHUNK; k += 256) { const int j = c0 + k; if (j < n) perm[j] = (int)(unsigned int)e[k]; }
        __syncthreads();
    }
}

// Shadow queue q of the sync-free loop (three of them, wave k uses k % 3): its ray counter and its fetch counter.
__host__ __device__ __forceinline__ int dt_cnt_shadow(int q) { return q == 0 ? DT_CNT_SHADOW : (q == 1 ? DT_CNT_SHADOW2 : DT_CNT_SHADOW3); }
__host__ __device__ __forceinline__ int dt_cnt_fetch_b(int q) { return q == 0 ? DT_CNT_FETCH_B : (q == 1 ? DT_CNT_FETCH_B2 : DT_CNT_FETCH_B3); }

// Device-side bookkeeping between two waves of the sync-free loop (one thread), after shade(k): the next wave's size, and the
// recycling of shadow queue `q_recycle` (= the queue wave k+1 will fill; last used by wave k-2, whose shadow pass is complete).
// A wave that overflowed a queue (DT_CNT_OVERFLOW set by dt_emit_child / dt_emit_shadow, or a counter past its capacity) ends the
// frame on the device: the next wave's size becomes 0, so every later launch of the frame is a no-op and nothing indexes past an
// allocation before the host sees the flag and retries with smaller waves.
__global__ void k_wave_advance(int* c, int q_recycle, int capacity, int shadow_capacity) {
    unsigned long long* tot_c = reinterpret_cast<unsigned long long*>(c + DT_CNT_TOT_CLOSEST);
    unsigned long long* tot_s = reinterpret_cast<unsigned long long*>(c + DT_CNT_TOT_SHADOW);
    const bool overflow = c[DT_CNT_OVERFLOW] != 0 || c[DT_CNT_NEXT] > capacity;
    if (overflow) c[DT_CNT_OVERFLOW] = 1;
    *tot_c += (unsigned long long)min(c[DT_CNT_NEXT], capacity);
    *tot_s += (unsigned long long)min(c[dt_cnt_shadow(q_recycle)], shadow_capacity);
    c[DT_CNT_CUR] = overflow ? 0 : c[DT_CNT_NEXT];
    c[DT_CNT_NEXT] = 0;
    c[DT_CNT_FETCH_A] = 0;
    c[dt_cnt_shadow(q_recycle)] = 0;
    c[dt_cnt_fetch_b(q_recycle)] = 0;
}

// Deferred mesh-light NEE (see k_shade): after the next wave's closest-hit pass, drop the entries whose GI
// child hit the emissive shape carrying the same light id; the survivors are traced like any shadow ray.
__global__ void k_filter_deferred(DtSceneDev S, DtShadowQueue sq, int n, DtRayQueue next) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int2 df = sq.defer[i];
    if (df.x < 0) return;
    const int hs = __float_as_int(next.hit0[df.x].w);
    if (hs < 0) return;
    const DtShapeDev& sh = S.shapes[hs];
    if (S.materials[sh.material - 1].type == DT_MAT_EMISSIVE && sh.id == df.y)
        sq.contrib_pix[i] = make_float4(0.f, 0.f, 0.f, sq.contrib_pix[i].w);
}

// ------------------------------------------------------------------ device-resident wave loop
// Path tracing with Russian roulette has no depth bound and frames of many samples need several batches, so the number of waves
// is only known on the device.  Instead of one host round trip per wave the whole frame is ONE CUDA graph: a WHILE node whose
// body is two waves (even: queue 0 -> 1, odd: 1 -> 0; kernel parameters are fixed inside a graph) followed by k_tail.  All wave
// sizes live in the counter block; k_loop_begin tops the wave up with new camera samples, k_loop_end advances to the next wave
// and (odd wave) decides through the graph's conditional handle whether the body runs again.
__device__ __forceinline__ bool dt_deferred_skipped(const DtSceneDev& S, const int2 df, const float4* __restrict__ child_hit0) {
    if (df.x < 0) return false;
    const int hs = __float_as_int(child_hit0[df.x].w);
    if (hs < 0) return false;
    const DtShapeDev& sh = S.shapes[hs];
    return S.materials[sh.material - 1].type == DT_MAT_EMISSIVE && sh.id == df.y;
}
__global__ void k_filter_deferred_dev(DtSceneDev S, DtShadowQueue sq, const int* n_ptr, int n_cap, DtRayQueue next) {
    const int n = min(*n_ptr, n_cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (dt_deferred_skipped(S, sq.defer[i], next.hit0)) sq.contrib_pix[i] = make_float4(0.f, 0.f, 0.f, sq.contrib_pix[i].w);
}

__global__ void __launch_bounds__(2 * DT_SORT_BINS) k_loop_begin(int* c, int wave_max, long long total, int* sort_hist) {
    if (sort_hist) sort_hist[threadIdx.x] = 0;                                   // bin counts + running cursors of this wave's sort stage
    if (threadIdx.x != 0) return;
    const int count = c[DT_CNT_CUR];
    long long* np = reinterpret_cast<long long*>(c + DT_CNT_NEXT_PRIMARY);
    int n_new = 0;
    if (count < wave_max && *np < total && c[DT_CNT_OVERFLOW] == 0) n_new = (int)min((long long)(wave_max - count), total - *np);
    c[DT_CNT_GEN_N] = n_new; c[DT_CNT_GEN_BASE] = count;
    *reinterpret_cast<long long*>(c + DT_CNT_GEN_K0) = *np;
    *np += n_new;
    c[DT_CNT_CUR] = count + n_new;
    c[DT_CNT_NEXT] = 0; c[DT_CNT_SHADOW] = 0; c[DT_CNT_FETCH_A] = 0; c[DT_CNT_FETCH_B] = 0;
}

// handle_valid: this is the odd wave of the WHILE body -> decide whether the body runs again.  The loop hands over to k_tail
// once every camera sample has been generated and at most tail_threshold rays are alive (tail_threshold < 0: never).
__global__ void k_loop_end(int* c, int capacity, int shadow_capacity, int defer, long long total, int tail_threshold, int handle_valid, cudaGraphConditionalHandle handle) {
    unsigned long long* tot_c = reinterpret_cast<unsigned long long*>(c + DT_CNT_TOT_CLOSEST);
    unsigned long long* tot_s = reinterpret_cast<unsigned long long*>(c + DT_CNT_TOT_SHADOW);
    const bool overflow = c[DT_CNT_OVERFLOW] != 0 || c[DT_CNT_NEXT] > capacity || c[DT_CNT_SHADOW] > shadow_capacity;
    if (overflow) c[DT_CNT_OVERFLOW] = 1;
    const int n_shadow = min(c[DT_CNT_SHADOW], shadow_capacity);
    if (c[DT_CNT_CUR] > 0 || c[DT_CNT_PREV_SHADOW] > 0) c[DT_CNT_WAVES]++;
    *tot_c += (unsigned long long)min(c[DT_CNT_NEXT], capacity);
    *tot_s += (unsigned long long)n_shadow;
    c[DT_CNT_PREV_SHADOW] = (defer && !overflow) ? n_shadow : 0;
    c[DT_CNT_CUR] = overflow ? 0 : c[DT_CNT_NEXT];
    if (handle_valid) {
        c[DT_CNT_ITERS]++;
        const long long np = *reinterpret_cast<const long long*>(c + DT_CNT_NEXT_PRIMARY);
        const bool more = !overflow && (c[DT_CNT_CUR] > 0 || np < total || c[DT_CNT_PREV_SHADOW] > 0);
        const bool tail = np >= total && c[DT_CNT_CUR] <= tail_threshold && c[DT_CNT_PREV_SHADOW] <= tail_threshold;
        cudaGraphSetConditional(handle, (more && !tail) ? 1u : 0u);
    }
}

// The last few thousand rays of a Russian-roulette frame live for thousands of bounces (the reference's RR never ends a pure GI
// chain, raytracer.cpp:137-147; on the 10 M-triangle mesh a few rays slip between two triangles into the closed mesh and bounce
// inside it).  A wave of a handful of rays is all latency, so the survivors are dealt out to the blocks of ONE kernel and every
// block runs the complete wave loop -- closest hit, deferred NEE, shade, shadow -- on block-private queues with barriers where
// the frame loop has kernel boundaries.  The bounce chain closest(i) -> shade(i) -> closest(i + 1) is the critical path (one
// dependent DRAM / L2 access after the other); the shadow rays are not on it, so the block is split into two PATH warps, which run
// that chain, and two SHADOW warps, which trace the shadow rays emitted two waves earlier (three rotating shadow buffers; the
// deferred mesh-light entries of wave i - 2 need the closest hits of wave i - 1, which stay intact while wave i is processed).
// The ray tree and the per-path RNG streams are the frame loop's, so the image does not depend on where the hand-over happens.
struct DtTailMem {
    DtRayQueue q[2];              // G x capacity entries each (block b owns [b * capacity, (b + 1) * capacity))
    float4* miss[2];
    DtShadowQueue sq;             // G x 3 x shadow_capacity (three rotating buffers per block)
    int capacity, shadow_capacity;
};
__device__ __forceinline__ DtRayQueue dt_queue_at(const DtRayQueue& q, size_t off) {
    DtRayQueue r;
    r.o_time = q.o_time + off; r.d_tmax = q.d_tmax + off; r.hit0 = q.hit0 + off; r.hit_face = q.hit_face + off; r.pixel = q.pixel + off;
    r.weight_n = q.weight_n + off; r.thr_beer = q.thr_beer + off; r.misc = q.misc + off; r.sort_key = q.sort_key ? q.sort_key + off : nullptr;
    return r;
}
__device__ __forceinline__ DtShadowQueue dt_shadow_queue_at(const DtShadowQueue& q, size_t off) {
    DtShadowQueue r;
    r.o_time = q.o_time + off; r.d_tmax = q.d_tmax + off; r.contrib_pix = q.contrib_pix + off; r.defer = q.defer ? q.defer + off : nullptr;
    return r;
}
#define DT_TAIL_PATH_THREADS 64
#ifndef DT_TAIL_DECOUPLED
#define DT_TAIL_DECOUPLED 1
#endif
#if DT_TAIL_DECOUPLED
// Decoupled variant (default).  In the version below (kept for A/B, -DDT_TAIL_DECOUPLED=0) the shadow warps must finish the shadow
// rays of wave i - 2 inside wave i, and they -- not the bounce chain -- were the longer side of most late waves.  Here every
// rotating shadow buffer has its own warp, and the block-wide barrier is replaced by named barriers that express the actual hazards:
//   FULL[k]   path -> shadow warp k: the shadow rays of batch b (emitted by shade(b), buffer k = b mod 3) are complete and their GI
//             children (wave b + 1) have their closest hits                                         (arrive after closest(b + 1))
//   CHECK[k]  shadow warp k -> path: the deferred mesh-light entries of batch b have looked at their child's hit record
//             (closest(b + 3) overwrites those records: same queue parity)                          (waited for before closest(b + 3))
//   DRAIN[k]  shadow warp k -> path: batch b is traced, buffer k may be refilled                   (waited for before shade(b + 3))
// so a batch has shade(b + 1) + wave b + 2 + closest(b + 3) to finish, about two waves instead of one, and the bounce chain of the
// two path warps no longer waits for shadow rays unless a batch takes twice as long as a wave.
#define DT_TAIL_THREADS (DT_TAIL_PATH_THREADS + 96)
#define DT_TAIL_MINBLOCKS 3                       // 3 x 160 threads at 128 registers
__device__ __forceinline__ void dt_bar_sync(int id, int n) { __syncwarp(); asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void dt_bar_arrive(int id, int n) { __syncwarp(); __threadfence_block(); asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(n) : "memory"); }
#define DT_BAR_PATH 1
#define DT_BAR_FULL 2
#define DT_BAR_CHECK 5
#define DT_BAR_DRAIN 8
template <bool STALE>
__global__ void __launch_bounds__(DT_TAIL_THREADS, DT_TAIL_MINBLOCKS) k_tail(DtSceneDev S, DtCamDev cam, DtRayQueue gq, const float4* gmiss, DtShadowQueue gsq, int* c, DtTailMem M, int defer, float4* accum) {
    // 0 rays of this wave, 1 rays emitted for the next wave, 2 overflow, 3..5 entries of the three shadow buffers, 6 / 7 untraced shadow / closest-hit entries
    __shared__ int sc[8];
    __shared__ int s_exit;
    const int b = blockIdx.x, G = gridDim.x, tid = threadIdx.x;
    const int n = c[DT_CNT_CUR], nps = defer ? c[DT_CNT_PREV_SHADOW] : 0;
    if ((n == 0 && nps == 0) || c[DT_CNT_OVERFLOW] != 0) return;
    const DtRayQueue L[2] = {dt_queue_at(M.q[0], (size_t)b * M.capacity), dt_queue_at(M.q[1], (size_t)b * M.capacity)};
    float4* const Lmiss[2] = {M.miss[0] ? M.miss[0] + (size_t)b * M.capacity : nullptr, M.miss[1] ? M.miss[1] + (size_t)b * M.capacity : nullptr};
    const DtShadowQueue B[3] = {dt_shadow_queue_at(M.sq, ((size_t)b * 3 + 0) * M.shadow_capacity), dt_shadow_queue_at(M.sq, ((size_t)b * 3 + 1) * M.shadow_capacity),
                                dt_shadow_queue_at(M.sq, ((size_t)b * 3 + 2) * M.shadow_capacity)};
    const int chunk = (n + G - 1) / G;
    const int lo = min(n, b * chunk), hi = min(n, lo + chunk);
    if (tid < 8) sc[tid] = 0;
    if (tid == 8) s_exit = 0;
    __syncthreads();
    for (int j = tid; j < hi - lo; j += blockDim.x) {
        const int g = lo + j;
        L[0].o_time[j] = gq.o_time[g]; L[0].d_tmax[j] = gq.d_tmax[g]; L[0].pixel[j] = gq.pixel[g];
        L[0].weight_n[j] = gq.weight_n[g]; L[0].thr_beer[j] = gq.thr_beer[g]; L[0].misc[j] = gq.misc[g];
        if (Lmiss[0] && gmiss) Lmiss[0][j] = gmiss[g];
    }
    if (tid == 0) { sc[0] = hi - lo; if (b == 0) c[DT_CNT_TAIL_RAYS] = n; }
    // The deferred NEE entries of the frame loop's last wave follow the block that owns their GI child; entries without a child
    // are dealt round-robin.  They are batch -1 (buffer 2): their children are the rays of tail wave 0.
    for (int e = tid; e < nps; e += blockDim.x) {
        const int2 df = gsq.defer[e];
        const bool mine = df.x >= 0 ? (df.x >= lo && df.x < hi) : (e % G == b);
        if (!mine || gsq.d_tmax[e].w < 0.0f) continue;                  // (untraced entries stay behind)
        const int slot = atomicAdd(&sc[5], 1);
        if (slot >= M.shadow_capacity) { sc[2] = 1; continue; }
        B[2].o_time[slot] = gsq.o_time[e]; B[2].d_tmax[slot] = gsq.d_tmax[e]; B[2].contrib_pix[slot] = gsq.contrib_pix[e];
        B[2].defer[slot] = make_int2(df.x >= 0 ? df.x - lo : -1, df.y);
    }
    __syncthreads();
    DT_DECLARE_STACK(stack);
    constexpr int PT = DT_TAIL_PATH_THREADS, PS = DT_TAIL_PATH_THREADS + 32;
    if (tid >= PT) {
        // ---- shadow warp k: batches k, k + 3, ... (warp 2 starts with the handed-over batch -1)
        const int k = (tid - PT) >> 5, lane = tid & 31;
        const DtShadowQueue& Q = B[k];
        for (int bi = (k == 2 ? -1 : k);; bi += 3) {
            dt_bar_sync(DT_BAR_FULL + k, PS);
            if (s_exit) break;
            const int pend = min(sc[3 + k], M.shadow_capacity);
            if (defer) {
                const float4* child_hit0 = L[(bi + 1) & 1].hit0;
                for (int e = lane; e < pend; e += 32)
                    if (dt_deferred_skipped(S, Q.defer[e], child_hit0)) Q.d_tmax[e].w = -1.0f;          // the child found the light itself: not traced
                __syncwarp();
            }
            dt_bar_arrive(DT_BAR_CHECK + k, PS);
            for (int e = lane; e < pend; e += 32) {
                const float4 d = Q.d_tmax[e];
                if (d.w < 0.0f) continue;                               // zero contribution (dt_shade_ray) or skipped above
                const float4 o = Q.o_time[e];
                DtTrav T;
                dt_trav_init<true>(T, S, V(o.x, o.y, o.z), V(d.x, d.y, d.z), o.w, d.w);
                while (!dt_trav_step<true, true>(T, stack, S, Q.o_time + e, Q.d_tmax + e)) {}
                dt_store_shadow(Q, e, T.best, accum);
            }
            __syncwarp();
            dt_bar_arrive(DT_BAR_DRAIN + k, PS);
        }
        return;
    }
    // ---- path warps: closest(i) -> shade(i) -> closest(i + 1) ...
    int waves = 0, i = 0;
    unsigned long long n_closest = 0, n_shadow = 0;          // thread 0 only
#ifdef DT_TAIL_PROFILE
    long long tp[3] = {0, 0, 0}, tp_last = clock64(), tp_rays = 0;
#define DT_TP(i) { const long long now_ = clock64(); tp[i] += now_ - tp_last; tp_last = now_; }
#else
#define DT_TP(i)
#endif
    for (;; i++) {
        const int cur = sc[0];
        // batches i - 1 (not announced yet) and i - 2 empty and no ray left: done (batch i - 3 may still be in flight, see the drain below)
        if ((cur == 0 && sc[3 + (i + 2) % 3] == 0 && sc[3 + (i + 1) % 3] == 0) || sc[2] != 0) break;
        const DtRayQueue& in = L[i & 1];
        if (i >= 2) dt_bar_sync(DT_BAR_CHECK + i % 3, PS);            // batch i - 3 has read the hit records of wave i - 2 (this queue)
        DT_TP(2)
        for (int j = tid; j < cur; j += PT) {
            if (in.pixel[j] == DT_DEAD_PIXEL) continue;
            const float4 o = in.o_time[j], d = in.d_tmax[j];
            DtTrav T;
            dt_trav_init<false>(T, S, V(o.x, o.y, o.z), V(d.x, d.y, d.z), o.w, CUDART_INF_F);
            while (!dt_trav_step<false, true>(T, stack, S, in.o_time + j, in.d_tmax + j)) {}
            dt_store_closest(in, j, T.best);
        }
        dt_bar_sync(DT_BAR_PATH, PT);                                  // all closest hits of the wave are stored
        dt_bar_arrive(DT_BAR_FULL + (i + 2) % 3, PS);                  // batch i - 1 may start
        DT_TP(0)
        if (i >= 2) dt_bar_sync(DT_BAR_DRAIN + i % 3, PS);            // batch i - 3 is traced: its buffer is free
        if (tid == 0) sc[3 + i % 3] = 0;
        dt_bar_sync(DT_BAR_PATH, PT);
        DT_TP(2)
        const DtShadeCounters cnt = {&sc[1], &sc[3 + i % 3], &sc[2], nullptr, nullptr};
        for (int j0 = tid & ~31; j0 < cur; j0 += PT) {                  // warp-uniform trip count, see k_shade
            const int j = j0 + (tid & 31);
            if (j < cur) dt_shade_ray<STALE>(j, S, cam, in, Lmiss[i & 1], L[(i + 1) & 1], Lmiss[(i + 1) & 1], M.capacity, B[i % 3], M.shadow_capacity, cnt, accum, &sc[6]);
            __syncwarp();
        }
        dt_bar_sync(DT_BAR_PATH, PT);
        DT_TP(1)
        if (tid == 0) {
            if (cur > 0) waves++;
            if (sc[1] > M.capacity || sc[3 + i % 3] > M.shadow_capacity) sc[2] = 1;
            n_closest += (unsigned long long)min(sc[1], M.capacity); n_shadow += (unsigned long long)min(sc[3 + i % 3], M.shadow_capacity);
            sc[0] = min(sc[1], M.capacity);
            sc[1] = 0;
#ifdef DT_TAIL_PROFILE
            tp_rays += cur;
#endif
        }
        dt_bar_sync(DT_BAR_PATH, PT);
    }
    // drain: the announced batches nobody has waited for yet (i - 3, i - 2), then release the three warps from their FULL barrier
    for (int bb = i - 3; bb <= i - 2; bb++)
        if (bb >= -1) { dt_bar_sync(DT_BAR_CHECK + (bb + 3) % 3, PS); dt_bar_sync(DT_BAR_DRAIN + (bb + 3) % 3, PS); }
    if (tid == 0) s_exit = 1;
    dt_bar_sync(DT_BAR_PATH, PT);
    for (int k = 0; k < 3; k++) dt_bar_arrive(DT_BAR_FULL + k, PS);
    if (tid == 0) {
        if (sc[6] > 0) atomicAdd(reinterpret_cast<unsigned long long*>(c + DT_CNT_SHADOW_DEAD), (unsigned long long)sc[6]);
        if (sc[7] > 0) atomicAdd(reinterpret_cast<unsigned long long*>(c + DT_CNT_CLOSEST_DEAD), (unsigned long long)sc[7]);
        atomicAdd(reinterpret_cast<unsigned long long*>(c + DT_CNT_TOT_CLOSEST), n_closest);
        atomicAdd(reinterpret_cast<unsigned long long*>(c + DT_CNT_TOT_SHADOW), n_shadow);
        atomicMax(c + DT_CNT_TAIL_WAVES, waves);
        if (sc[2] != 0) atomicAdd(c + DT_CNT_OVERFLOW, 1);
#ifdef DT_TAIL_PROFILE
        if (waves > 1000) printf("[dt-tail] block %d: %d waves, %lld ray-waves | cycles per wave: closest %lld, shade %lld, waiting for shadow warps + bookkeeping %lld\n",
                                 b, waves, tp_rays, tp[0] / waves, tp[1] / waves, tp[2] / waves);
#endif
    }
#undef DT_TP
}
#else
#define DT_TAIL_THREADS 128
template <bool STALE>
__global__ void __launch_bounds__(128, DT_SHADE_MINBLOCKS) k_tail(DtSceneDev S, DtCamDev cam, DtRayQueue gq, const float4* gmiss, DtShadowQueue gsq, int* c, DtTailMem M, int defer, float4* accum) {
    // 0 rays of this wave, 1 rays emitted for the next wave, 2 overflow, 3..5 entries of the three shadow buffers, 6 / 7 untraced shadow / closest-hit entries
    __shared__ int sc[8];
#ifdef DT_TAIL_PROFILE
    __shared__ long long tp_shadow;
#endif
    const int b = blockIdx.x, G = gridDim.x, tid = threadIdx.x;
    const int n = c[DT_CNT_CUR], nps = defer ? c[DT_CNT_PREV_SHADOW] : 0;
    if ((n == 0 && nps == 0) || c[DT_CNT_OVERFLOW] != 0) return;
    const DtRayQueue L[2] = {dt_queue_at(M.q[0], (size_t)b * M.capacity), dt_queue_at(M.q[1], (size_t)b * M.capacity)};
    float4* const Lmiss[2] = {M.miss[0] ? M.miss[0] + (size_t)b * M.capacity : nullptr, M.miss[1] ? M.miss[1] + (size_t)b * M.capacity : nullptr};
    const DtShadowQueue B[3] = {dt_shadow_queue_at(M.sq, ((size_t)b * 3 + 0) * M.shadow_capacity), dt_shadow_queue_at(M.sq, ((size_t)b * 3 + 1) * M.shadow_capacity),
                                dt_shadow_queue_at(M.sq, ((size_t)b * 3 + 2) * M.shadow_capacity)};
    const int chunk = (n + G - 1) / G;
    const int lo = min(n, b * chunk), hi = min(n, lo + chunk);
    if (tid < 8) sc[tid] = 0;
#ifdef DT_TAIL_PROFILE
    if (tid == 0) tp_shadow = 0;
#endif
    __syncthreads();
    for (int j = tid; j < hi - lo; j += blockDim.x) {
        const int g = lo + j;
        L[0].o_time[j] = gq.o_time[g]; L[0].d_tmax[j] = gq.d_tmax[g]; L[0].pixel[j] = gq.pixel[g];
        L[0].weight_n[j] = gq.weight_n[g]; L[0].thr_beer[j] = gq.thr_beer[g]; L[0].misc[j] = gq.misc[g];
        if (Lmiss[0] && gmiss) Lmiss[0][j] = gmiss[g];
    }
    if (tid == 0) { sc[0] = hi - lo; if (b == 0) c[DT_CNT_TAIL_RAYS] = n; }
    // The deferred NEE entries of the frame loop's last wave follow the block that owns their GI child; entries without a child
    // are dealt round-robin.  They enter as "emitted one wave ago" (buffer 2), i.e. they are traced during the second tail wave.
    for (int e = tid; e < nps; e += blockDim.x) {
        const int2 df = gsq.defer[e];
        const bool mine = df.x >= 0 ? (df.x >= lo && df.x < hi) : (e % G == b);
        if (!mine || gsq.d_tmax[e].w < 0.0f) continue;                  // (untraced entries stay behind)
        const int slot = atomicAdd(&sc[5], 1);
        if (slot >= M.shadow_capacity) { sc[2] = 1; continue; }
        B[2].o_time[slot] = gsq.o_time[e]; B[2].d_tmax[slot] = gsq.d_tmax[e]; B[2].contrib_pix[slot] = gsq.contrib_pix[e];
        B[2].defer[slot] = make_int2(df.x >= 0 ? df.x - lo : -1, df.y);
    }
    __syncthreads();
    int waves = 0;
    unsigned long long n_closest = 0, n_shadow = 0;          // thread 0 only
    DT_DECLARE_STACK(stack);
#ifdef DT_TAIL_PROFILE
    long long tp[3] = {0, 0, 0}, tp_last = clock64(), tp_rays = 0;
#define DT_TP(i) { const long long now_ = clock64(); tp[i] += now_ - tp_last; tp_last = now_; }
#else
#define DT_TP(i)
#endif
    for (int i = 0;; i++) {
        const int cur = sc[0];
        const int jb = (i + 1) % 3;                               // shadow buffer filled during wave i - 2
        const int pend = min(sc[3 + jb], M.shadow_capacity);
        if ((cur == 0 && sc[3] == 0 && sc[4] == 0 && sc[5] == 0) || sc[2] != 0) break;
        const DtRayQueue& in = L[i & 1];
        if (tid < DT_TAIL_PATH_THREADS) {
            // ---- path warps: closest(i) -> shade(i)
            for (int j = tid; j < cur; j += DT_TAIL_PATH_THREADS) {
                if (in.pixel[j] == DT_DEAD_PIXEL) continue;
                const float4 o = in.o_time[j], d = in.d_tmax[j];
                DtTrav T;
                dt_trav_init<false>(T, S, V(o.x, o.y, o.z), V(d.x, d.y, d.z), o.w, CUDART_INF_F);
                while (!dt_trav_step<false, true>(T, stack, S, in.o_time + j, in.d_tmax + j)) {}
                dt_store_closest(in, j, T.best);
            }
            asm volatile("bar.sync 1, %0;" :: "n"(DT_TAIL_PATH_THREADS) : "memory");          // all closest hits of the wave are stored (GI children read none, but shade reads hit0 by index)
            DT_TP(0)
            const DtShadeCounters cnt = {&sc[1], &sc[3 + i % 3], &sc[2], nullptr, nullptr};
            for (int j0 = tid & ~31; j0 < cur; j0 += DT_TAIL_PATH_THREADS) {                  // warp-uniform trip count, see k_shade
                const int j = j0 + (tid & 31);
                if (j < cur) dt_shade_ray<STALE>(j, S, cam, in, Lmiss[i & 1], L[(i + 1) & 1], Lmiss[(i + 1) & 1], M.capacity, B[i % 3], M.shadow_capacity, cnt, accum, &sc[6]);
                __syncwarp();
            }
            DT_TP(1)
        } else if (pend > 0) {
            // ---- shadow warps: the shadow rays emitted by shade(i - 2); deferred mesh-light entries look at the closest hit of
            // their GI child, a ray of wave i - 1 (queue (i - 1) & 1: its hit records are not rewritten before closest(i + 1))
#ifdef DT_TAIL_PROFILE
            const long long t0 = clock64();
#endif
            const DtShadowQueue& Q = B[jb];
            const float4* child_hit0 = L[(i + 1) & 1].hit0;
            for (int e = tid - DT_TAIL_PATH_THREADS; e < pend; e += blockDim.x - DT_TAIL_PATH_THREADS) {
                if (defer && dt_deferred_skipped(S, Q.defer[e], child_hit0)) continue;
                const float4 o = Q.o_time[e], d = Q.d_tmax[e];
                if (d.w < 0.0f) continue;                               // zero contribution: not traced (dt_shade_ray)
                DtTrav T;
                dt_trav_init<true>(T, S, V(o.x, o.y, o.z), V(d.x, d.y, d.z), o.w, d.w);
                while (!dt_trav_step<true, true>(T, stack, S, Q.o_time + e, Q.d_tmax + e)) {}
                dt_store_shadow(Q, e, T.best, accum);
            }
#ifdef DT_TAIL_PROFILE
            if (tid == DT_TAIL_PATH_THREADS) tp_shadow += clock64() - t0;
#endif
        }
        __syncthreads();
        if (tid == 0) {
            if (cur > 0) waves++;
            if (sc[1] > M.capacity || sc[3 + i % 3] > M.shadow_capacity) sc[2] = 1;
            n_closest += (unsigned long long)min(sc[1], M.capacity); n_shadow += (unsigned long long)min(sc[3 + i % 3], M.shadow_capacity);
            sc[0] = min(sc[1], M.capacity);
            sc[1] = 0;
            sc[3 + jb] = 0;                                       // traced; shade(i + 1) fills this buffer next
#ifdef DT_TAIL_PROFILE
            tp_rays += cur;
#endif
        }
        __syncthreads();
        DT_TP(2)
    }
    if (tid == 0) {
        if (sc[6] > 0) atomicAdd(reinterpret_cast<unsigned long long*>(c + DT_CNT_SHADOW_DEAD), (unsigned long long)sc[6]);
        if (sc[7] > 0) atomicAdd(reinterpret_cast<unsigned long long*>(c + DT_CNT_CLOSEST_DEAD), (unsigned long long)sc[7]);
        atomicAdd(reinterpret_cast<unsigned long long*>(c + DT_CNT_TOT_CLOSEST), n_closest);
        atomicAdd(reinterpret_cast<unsigned long long*>(c + DT_CNT_TOT_SHADOW), n_shadow);
        atomicMax(c + DT_CNT_TAIL_WAVES, waves);
        if (sc[2] != 0) atomicAdd(c + DT_CNT_OVERFLOW, 1);
#ifdef DT_TAIL_PROFILE
        if (waves > 1000) printf("[dt-tail] block %d: %d waves, %lld ray-waves | cycles per wave: closest %lld, shade %lld, barrier + bookkeeping %lld | shadow warps %lld\n",
                                 b, waves, tp_rays, tp[0] / waves, tp[1] / waves, tp[2] / waves, tp_shadow / waves);
#endif
    }
#undef DT_TP
}
#endif

// ------------------------------------------------------------------ resolve
// main.cpp:97-125: Gaussian-weighted resolve, HDR store, LDR clamp((int)c) with cvttss2si semantics
__device__ __forceinline__ int dt_clamp_channel(float f) {
    int v;
    if (!(f > -2147483904.0f && f < 2147483648.0f)) v = (int)0x80000000; else v = (int)f;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}
__global__ void k_resolve(const float4* accum, int n_pix, int spp, float* hdr, uint8_t* ldr, int* counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const float4 a = accum[i];
    float r = a.x, g = a.y, b = a.z;
    if (spp > 1 && a.w > 0.f) { r = r / a.w; g = g / a.w; b = b / a.w; }
    if (isnan(r) || isnan(g) || isnan(b)) atomicAdd(counters + DT_CNT_NAN, 1);
    if (hdr) { hdr[3 * (size_t)i] = r; hdr[3 * (size_t)i + 1] = g; hdr[3 * (size_t)i + 2] = b; }
    if (ldr) {
        ldr[3 * (size_t)i] = (uint8_t)dt_clamp_channel(r);
        ldr[3 * (size_t)i + 1] = (uint8_t)dt_clamp_channel(g);
        ldr[3 * (size_t)i + 2] = (uint8_t)dt_clamp_channel(b);
    }
}

// Resolve + gather fused (DT_FLAG_PEER_FRAME): hdr / ldr may point into another GPU's memory (CUDA IPC mapping, stores travel
// over NVLink), so only owned pixels are written and nothing is read back from them.  One warp per owned strip of
// DT_TILE_GROUP tiles (64x4 pixels): a lane resolves 4 consecutive pixels of two rows; the 12 LDR bytes it produces are
// exchanged by shuffles so that every store instruction writes whole 32-bit words, 64 contiguous bytes per row
// (fast path: full strip, width a multiple of 4 so that rows are word-aligned; otherwise byte stores).
__global__ void k_resolve_tiles(const float4* accum, int width, int height, int tiles_x, int tiles_y, long long my_tiles, int tile_rank, int tile_world,
                                int spp, float* hdr, uint8_t* ldr, int* counters) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;      // strip slot of this rank
    if (w * DT_TILE_GROUP >= my_tiles) return;
    const int lane = threadIdx.x & 31;
    int tx0, ty;
    if (!dt_rank_tile(w * DT_TILE_GROUP, tile_rank, tile_world, tiles_x, tiles_y, tx0, ty)) return;
    const int x0 = tx0 * 8, y0 = ty * 4;
    const int strip_w = min(DT_TILE_GROUP * 8, min(tiles_x * 8, width) - x0);      // pixels of this strip inside the image
    const bool fast = ldr && (width & 3) == 0 && strip_w == DT_TILE_GROUP * 8;
    const int l16 = lane & 15;
    for (int half = 0; half < 2; half++) {
        const int y = y0 + half * 2 + (lane >> 4);
        uint32_t words[3] = {0u, 0u, 0u};
        uint8_t* bytes = reinterpret_cast<uint8_t*>(words);
        for (int k = 0; k < 4; k++) {
            const int x = x0 + l16 * 4 + k;
            if (y >= height || x >= x0 + strip_w) continue;
            const size_t i = (size_t)y * width + x;
            const float4 a = accum[i];
            float r = a.x, g = a.y, b = a.z;
            if (spp > 1 && a.w > 0.f) { r = r / a.w; g = g / a.w; b = b / a.w; }
            if (isnan(r) || isnan(g) || isnan(b)) atomicAdd(counters + DT_CNT_NAN, 1);
            if (hdr) { hdr[3 * i] = r; hdr[3 * i + 1] = g; hdr[3 * i + 2] = b; }
            if (ldr) {
                const uint8_t cr = (uint8_t)dt_clamp_channel(r), cg = (uint8_t)dt_clamp_channel(g), cb = (uint8_t)dt_clamp_channel(b);
                if (fast) { bytes[3 * k] = cr; bytes[3 * k + 1] = cg; bytes[3 * k + 2] = cb; }
                else { ldr[3 * i] = cr; ldr[3 * i + 1] = cg; ldr[3 * i + 2] = cb; }
            }
        }
        if (fast) {                                                                  // warp-uniform
            // a row of the strip is 48 words; lane l16 holds words 3*l16 .. 3*l16+2 and stores words l16, 16 + l16, 32 + l16
            uint32_t* row = reinterpret_cast<uint32_t*>(ldr + ((size_t)y * width + x0) * 3);
            for (int k = 0; k < 3; k++) {
                const int j = 16 * k + l16, src = (lane & 16) | (j / 3), slot = j % 3;
                const uint32_t w0 = __shfl_sync(0xFFFFFFFFu, words[0], src), w1 = __shfl_sync(0xFFFFFFFFu, words[1], src), w2 = __shfl_sync(0xFFFFFFFFu, words[2], src);
                if (y < height) row[j] = slot == 0 ? w0 : (slot == 1 ? w1 : w2);
            }
        }
    }
}

// LDR clamp of an already resolved radiance buffer (main.cpp:118-125)
__global__ void k_clamp_hdr(const float* hdr, int n_pix, uint8_t* ldr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    ldr[3 * (size_t)i] = (uint8_t)dt_clamp_channel(hdr[3 * (size_t)i]);
    ldr[3 * (size_t)i + 1] = (uint8_t)dt_clamp_channel(hdr[3 * (size_t)i + 1]);
    ldr[3 * (size_t)i + 2] = (uint8_t)dt_clamp_channel(hdr[3 * (size_t)i + 2]);
}

// ------------------------------------------------------------------ tonemap (tonemapper.h:28-119)
// pass 1: sum of log(delta + Y) in double;  pass 2: exact k-th smallest of all 3N channel values by 4-pass
// radix select on order-preserving keys (the reference sorts all 3N floats, tonemapper.h:51);  pass 3: map.
__device__ __forceinline__ uint32_t dt_float_key(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float dt_key_float(uint32_t k) { uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k; return __uint_as_float(u); }

__global__ void k_tm_logsum(const float* hdr, int n_pix, double* sum) {
    double local = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += gridDim.x * blockDim.x) {
        double r = hdr[3 * (size_t)i], g = hdr[3 * (size_t)i + 1], b = hdr[3 * (size_t)i + 2];
        double lum = 0.2126 * r + 0.7152 * g + 0.0722 * b;
        local += log((double)0.01f + lum);
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xFFFFFFFFu, local, o);
    __shared__ double ws[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) ws[wid] = local;
    __syncthreads();
    if (wid == 0) {
        local = lane < (blockDim.x >> 5) ? ws[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xFFFFFFFFu, local, o);
        if (lane == 0) atomicAdd(sum, local);
    }
}
// histogram of byte `shift/8` of the keys whose higher bytes equal `prefix`
__global__ void k_tm_hist(const float* vals, size_t n, const uint32_t* prefix_ptr, uint32_t prefix_mask, int shift, unsigned int* hist) {
    __shared__ unsigned int sh[256];
    const uint32_t prefix = *prefix_ptr;                  // refined by k_tm_pick of the previous pass: no host round trip between the passes
    for (int k = threadIdx.x; k < 256; k += blockDim.x) sh[k] = 0;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t key = dt_float_key(vals[i]);
        if ((key & prefix_mask) == prefix) atomicAdd(&sh[(key >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 256; k += blockDim.x) if (sh[k]) atomicAdd(&hist[k], sh[k]);
}
// single-thread step of the select: pick the bucket holding rank `*rank`, refine prefix
__global__ void k_tm_pick(unsigned int* hist, unsigned long long* rank, uint32_t* prefix, int shift) {
    unsigned long long r = *rank;
    uint32_t b = 0;
    for (; b < 256; b++) { unsigned int c = hist[b]; if (r < c) break; r -= c; }
    if (b > 255) b = 255;
    *rank = r;
    *prefix |= (b << shift);
    for (int k = 0; k < 256; k++) hist[k] = 0;
}
__global__ void k_tm_map(const float* hdr, int n_pix, const double* logsum, const uint32_t* white_key, float key, float burn,
                         float saturation, float gamma, uint8_t* ldr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const double avgLum = exp(*logsum / (double)n_pix);
    double R = hdr[3 * (size_t)i], G = hdr[3 * (size_t)i + 1], B = hdr[3 * (size_t)i + 2];
    double y_i = 0.2126 * R + 0.7152 * G + 0.0722 * B;
    float y_of;
    {
        double Lxy = ((double)key * y_i) / avgLum;
        if (burn > 0.01) {
            double thr = (double)dt_key_float(*white_key);
            thr = thr * (double)key / avgLum;
            double LwhiteSqr = thr * thr;
            y_of = (float)((Lxy * (1 + (Lxy / LwhiteSqr))) / (1.0f + Lxy));
        } else y_of = (float)(Lxy / (1 + Lxy));
    }
    double y_o = y_of;
    double r_o = clipf((float)(y_o * pow((R / y_i), (double)saturation)), 0.0f, 1.0f);
    double g_o = clipf((float)(y_o * pow((G / y_i), (double)saturation)), 0.0f, 1.0f);
    double b_o = clipf((float)(y_o * pow((B / y_i), (double)saturation)), 0.0f, 1.0f);
    double gammaInv = 1.0f / gamma;
    int cr = (int)floor(fmin(255.0, 255 * pow(r_o, gammaInv)));
    int cg = (int)floor(fmin(255.0, 255 * pow(g_o, gammaInv)));
    int cb = (int)floor(fmin(255.0, 255 * pow(b_o, gammaInv)));
    ldr[3 * (size_t)i] = (uint8_t)cr; ldr[3 * (size_t)i + 1] = (uint8_t)cg; ldr[3 * (size_t)i + 2] = (uint8_t)cb;
}

// ------------------------------------------------------------------ small utility kernels
__global__ void k_pack_rays(const float* origins, const float* dirs, const float* tmax, int n, float4* o_time, float4* d_tmax) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    o_time[i] = make_float4(origins[3 * (size_t)i], origins[3 * (size_t)i + 1], origins[3 * (size_t)i + 2], 0.f);
    d_tmax[i] = make_float4(dirs[3 * (size_t)i], dirs[3 * (size_t)i + 1], dirs[3 * (size_t)i + 2], tmax ? tmax[i] : CUDART_INF_F);
}
__global__ void k_unpack_hits(const float4* hit0, const int32_t* hit_face, const uint32_t* pixel, int n, int32_t* shape, int32_t* face, float* t, int by_pixel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int dst = i;
    if (by_pixel) { if (pixel[i] == DT_DEAD_PIXEL) return; dst = (int)pixel[i]; }
    const float4 h = hit0[i];
    const int s = __float_as_int(h.w);
    shape[dst] = s; face[dst] = s >= 0 ? hit_face[i] : -1; t[dst] = s >= 0 ? h.x : CUDART_INF_F;
}
