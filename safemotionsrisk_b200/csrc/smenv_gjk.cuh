// smenv_gjk.cuh -- the narrow phase of the env step: ONE THREAD PER CONVEX PAIR QUERY.
//
// The planning kernels (smenv_plan.cuh) turn every env step into a flat list of 64-byte work items, one per convex
// pair that bounding volumes could not decide: "distance of robot shape A to obstacle shape B at the new knot"
// (getClosestPoints call sites ctlp.py:3267, :3300, :3353) or "do A and B touch in sub-step k" (getContactPoints,
// ctlp.py:4188, :4572).  gjk_kernel runs GJK on 32 different items per warp: the simplex arithmetic, which every
// lane of the former warp-per-pair version repeated redundantly, now does 32 times the work per instruction, and a
// support query scans only the ~20 candidates its direction's table cell lists (smenv_device.cuh, SM_LUT_RES)
// instead of all 252 vertices of an iiwa link.
//
// Exact pruning survives the parallelism: items of the same (env, distance class) sit next to each other in the
// list; the lanes that hold them form a group (match.any) and exchange, every iteration, the best upper bound any
// of them has reached (redux.min).  A lane whose separating-plane lower bound passes that value stops -- its pair
// cannot hold the minimum.  A finished pair reports its last upper bound with atomicMin on the env's result slot
// (the exact distance if it converged), and a newly loaded pair starts from what the slot already holds, so the
// slot ends up with the exact minimum over the pairs however the items are split over lanes, warps and time.
//
// Lanes finish at very different times (most pairs are pruned after one or two iterations, the closest pair of an
// env runs to convergence), so a warp refills its idle lanes from its chunk of the list while the others keep
// iterating: the first version, which ran 32 items to completion per warp, averaged 8.8 active lanes.
#pragma once
#include "smenv_geom.cuh"

enum { GJK_STATIC = 0, GJK_SELF = 1, GJK_MOVING = 2, GJK_CONTACT = 3, GJK_BRAKE = 4 };
// GJK_BRAKE: "is this pair closer than the safety distance at pose k of the human's braking trajectory" (ctlp.py:3155-3207);
// a touch test like GJK_CONTACT, reported as the first such pose in result slot 4

struct __align__(16) GjkItem {
    int env;
    uint32_t shapes;  // shape A | shape B << 16
    uint32_t meta;    // class | (1-based sub-step of a contact item) << 8
    float thr;        // distance classes: only distances <= thr count (query distance, tightened by the best upper
                      // bound known at planning time); contact class: the manifold's contact threshold
    float R[9];       // pose of B in the frame of A:  x_A = R x_B + t
    float t[3];
};
static_assert(sizeof(GjkItem) == 64, "GjkItem must stay 64 bytes");

#define SM_RES_STRIDE 8          /* per-env result record: static, self, moving distance keys, first contact sub-step,
                                    first colliding pose of the human's braking trajectory, 3 spare */
#define SM_RES_NO_CONTACT 0x7fffffffu

struct GjkArgs {
    const GjkItem* items;
    const int* n_items;  // device counter written by the planning kernels
    int capacity;
    unsigned* res;       // [n][SM_RES_STRIDE]
    unsigned long long* counters;
};

__device__ __forceinline__ float funkey(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// shared memory of gjk_kernel: hull vertices | direction tables | shapes
static_assert(sizeof(DevShape) % 16 == 0, "DevShape is copied in 16-byte vectors");
struct GjkSmem {
    float4* verts;
    uint32_t* lut;
    DevShape* shapes;
};
__device__ __forceinline__ GjkSmem gjk_carve(unsigned char* raw) {
    GjkSmem G;
    G.verts = reinterpret_cast<float4*>(raw);
    size_t off = (size_t)c_sc.n_verts * sizeof(float4);
    G.lut = reinterpret_cast<uint32_t*>(raw + off);
    off += (((size_t)c_sc.n_lut_words * 4) + 15) & ~(size_t)15;
    G.shapes = reinterpret_cast<DevShape*>(raw + off);
    return G;
}
static size_t gjk_smem_bytes(int n_verts, int n_lut_words, int n_shapes) {
    return (size_t)n_verts * sizeof(float4) + ((((size_t)n_lut_words * 4) + 15) & ~(size_t)15) +
           (size_t)n_shapes * sizeof(DevShape);
}

// support vertex of a hull in (local) direction d, scanned by one thread
__device__ __forceinline__ int support_thread(const float4* __restrict__ v, int n, const uint32_t* __restrict__ lut,
                                              int lut_off, V3 d, unsigned& ndots) {
    float best = -FLT_MAX;
    int bi = 0;
    if (lut_off >= 0) {
        const uint32_t* L = lut + (lut_off & SM_LUT_OFF_MASK);
        const uint32_t e = L[lut_cell_res(d.x, d.y, d.z, lut_res_of(lut_off))];
        const uint32_t* w = L + (e >> 8);
        const int cnt = (int)(e & 255u);
        ndots += (unsigned)cnt;
#pragma unroll 1
        for (int c = 0; c < cnt; c += 4) {
            const uint32_t word = *w++;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = (int)((word >> (8 * k)) & 255u);
                const float4 p = v[i];
                const float s = fmaf(p.x, d.x, fmaf(p.y, d.y, p.z * d.z));
                if (s > best) { best = s; bi = i; }
            }
        }
    } else {
        ndots += (unsigned)n;
#pragma unroll 2
        for (int i = 0; i < n; ++i) {
            const float4 p = v[i];
            const float s = fmaf(p.x, d.x, fmaf(p.y, d.y, p.z * d.z));
            if (s > best) { best = s; bi = i; }
        }
    }
    return bi;
}

#ifndef GJK_REFILL_MIN
#define GJK_REFILL_MIN 8 /* idle lanes of a warp before it fetches new items */
#endif
#ifndef GJK_CTA_CURSOR
#define GJK_CTA_CURSOR 1 /* warps draw items from their CTA's chunk (0: a fixed chunk per warp) */
#endif
#ifndef GJK_MIN_BLOCKS
#define GJK_MIN_BLOCKS 2 /* resident CTAs per SM the register allocation aims for */
#endif
#ifndef GJK_THREADS
#define GJK_THREADS 256  /* threads per CTA */
#endif
#ifndef GJK_FIRST_PRUNE
#define GJK_FIRST_PRUNE 1 /* the separating-plane bound is tested on the first direction (the line of centres) as well */
#endif

// THREADS: 256 with two CTAs per SM when the scene's shared-memory image (vertices + direction tables + shapes) allows
// it (sixteen warps per SM at up to 128 registers), else one CTA per SM of 768 threads (24 warps, 80 registers; Human
// scene: 4000 vertices; 512 and 1024 threads are kept for the sweep of tools/gjk_config_sweep.py)
template <bool COUNT, int THREADS = GJK_THREADS>
__global__ void __launch_bounds__(THREADS, THREADS <= 384 ? GJK_MIN_BLOCKS : 1) gjk_kernel(GjkArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int n_items = *A.n_items;
    if (n_items > A.capacity) n_items = A.capacity;
    // static partition of the list over the CTAs; inside a CTA the warps draw from the CTA's chunk through a cursor in
    // shared memory (no global cursor to contend on; the items of an env stay together, so that lanes of a warp hold
    // pairs that can prune each other; the warps of a CTA finish together instead of each draining its own tail)
    const int tid = threadIdx.x, lane = tid & 31;
#if GJK_CTA_CURSOR
    const int chunk = (n_items + (int)gridDim.x - 1) / (int)gridDim.x;
    if ((long long)blockIdx.x * chunk >= n_items) return;  // nothing for this block: no staging
    __shared__ int cta_cursor;
    if (tid == 0) cta_cursor = blockIdx.x * chunk;
    const int end = (blockIdx.x + 1) * chunk < n_items ? (blockIdx.x + 1) * chunk : n_items;
#else
    const int n_warps = gridDim.x * (blockDim.x >> 5);
    const int chunk = (n_items + n_warps - 1) / n_warps;
    if ((long long)blockIdx.x * (blockDim.x >> 5) * chunk >= n_items) return;  // nothing for this block: no staging
#endif
    GjkSmem G = gjk_carve(smem_raw);
    // hulls, direction tables and shapes -> shared memory: three bulk TMA copies issued by one thread
    __shared__ __align__(8) uint64_t stage_bar;
    if (tid == 0) {
        mbar_init(&stage_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        const uint32_t b_verts = (uint32_t)c_sc.n_verts * 16u, b_lut = (uint32_t)c_sc.n_lut_words * 4u,
                       b_shapes = (uint32_t)c_sc.n_shapes * (uint32_t)sizeof(DevShape);
        mbar_expect_tx(&stage_bar, b_verts + b_lut + b_shapes);
        tma_bulk_g2s(G.verts, c_sc.verts, b_verts, &stage_bar);
        tma_bulk_g2s(G.lut, c_sc.lut, b_lut, &stage_bar);
        tma_bulk_g2s(G.shapes, c_sc.scene_img, b_shapes, &stage_bar);   // SceneImage starts with the shapes
    }
    __syncthreads();                 // the barrier init is visible to everyone
    mbar_wait(&stage_bar, 0);
#if GJK_CTA_CURSOR
    bool more = true;   // warp-uniform: the CTA's chunk still held items at the warp's last draw
#else
    int cursor = (blockIdx.x * (blockDim.x >> 5) + (tid >> 5)) * chunk;
    const int end = cursor + chunk < n_items ? cursor + chunk : n_items;
#endif
    unsigned c_iters = 0, c_dots = 0, c_calls = 0;

    // ---------------- per-lane state of the pair in flight
    bool busy = false, have_point = false;
    int env = 0, cls = GJK_CONTACT, it = 0, nA = 0, nB = 0, lutA = -1, lutB = -1;
    unsigned meta = 0;
    float thr = 0.f, m = 0.f, lim_fixed = 0.f, lim = 0.f, touch = -1.f, vv = 0.f;
    float r0 = 1.f, r1 = 0.f, r2 = 0.f, r3 = 0.f, r4 = 1.f, r5 = 0.f, r6 = 0.f, r7 = 0.f, r8 = 1.f, tx = 0.f, ty = 0.f, tz = 0.f;
    const float4* vA = G.verts;
    const float4* vB = G.verts;
    V3 v = mk(1.f, 0.f, 0.f);
    Simplex S;
    S.n = 0;
    S.i0 = S.i1 = S.i2 = S.i3 = -1;
    S.p0 = S.p1 = S.p2 = S.p3 = mk(0.f, 0.f, 0.f);
#pragma unroll 1
    while (true) {
        // ---------------- refill: idle lanes take the next items of the warp's chunk
        const unsigned idle = __ballot_sync(FULL, !busy);
#if GJK_CTA_CURSOR
        int my = end;
        if (more && (idle == FULL || __popc(idle) >= GJK_REFILL_MIN)) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&cta_cursor, __popc(idle));
            base = __shfl_sync(FULL, base, 0);
            more = base + __popc(idle) < end;
            my = base + __popc(idle & ((1u << lane) - 1u));
        }
        {
#else
        if (cursor < end && (idle == FULL || __popc(idle) >= GJK_REFILL_MIN)) {
            const int my = cursor + __popc(idle & ((1u << lane) - 1u));
            cursor += __popc(idle);
#endif
            if (!busy && my < end) {
                const float4* ip = reinterpret_cast<const float4*>(A.items + my);
                const float4 h0 = __ldg(ip), h1 = __ldg(ip + 1), h2 = __ldg(ip + 2), h3 = __ldg(ip + 3);
                env = __float_as_int(h0.x);
                const unsigned shp = __float_as_uint(h0.y);
                meta = __float_as_uint(h0.z);
                thr = h0.w;
                cls = (int)(meta & 255u);
                r0 = h1.x; r1 = h1.y; r2 = h1.z; r3 = h1.w; r4 = h2.x; r5 = h2.y; r6 = h2.z; r7 = h2.w; r8 = h3.x;
                tx = h3.y; ty = h3.z; tz = h3.w;
                const DevShape& SA = G.shapes[shp & 0xffffu];
                const DevShape& SB = G.shapes[shp >> 16];
                vA = G.verts + SA.off; vB = G.verts + SB.off;
                nA = SA.cnt; nB = SB.cnt; lutA = SA.lut; lutB = SB.lut;
                m = SA.margin + SB.margin;
                lim_fixed = (cls >= GJK_CONTACT ? thr + 1e-3f : thr) + m;  // on the distance between the cores
                touch = cls >= GJK_CONTACT ? thr + m : -1.0f;
                lim = lim_fixed;
                if (cls < GJK_CONTACT) {  // what earlier pairs of the env already reached
                    const unsigned prev = __ldcg(A.res + (size_t)env * SM_RES_STRIDE + cls);
                    lim = fminf(lim_fixed, funkey(prev) + m);
                }
                // first direction: between the bounding-sphere centres (in the frame of A)
                v = mk(SA.cx - fmaf(r0, SB.cx, fmaf(r1, SB.cy, fmaf(r2, SB.cz, tx))),
                       SA.cy - fmaf(r3, SB.cx, fmaf(r4, SB.cy, fmaf(r5, SB.cz, ty))),
                       SA.cz - fmaf(r6, SB.cx, fmaf(r7, SB.cy, fmaf(r8, SB.cz, tz))));
                vv = dot(v, v);
                if (vv < 1e-12f) { v = mk(1.f, 0.f, 0.f); vv = 1.f; }
                S.n = 0;
                S.i0 = S.i1 = S.i2 = S.i3 = -1;
                have_point = false;
                it = 0;
                busy = true;
                if (COUNT) c_calls++;
            }
        }
        if (!__any_sync(FULL, busy)) break;
        // lanes with the same (env, distance class) prune each other; contact items and idle lanes share one dummy
        // group (redux over lane-dependent masks is serialised per distinct mask)
        const bool dist = busy && cls < GJK_CONTACT;
        const unsigned peers = __match_any_sync(FULL, dist ? (((unsigned)env << 2) | (unsigned)cls) : 0xffffffffu);
        bool done = false;
        if (busy) {
            // ---------------- one GJK iteration in the frame of A
            const int sa = support_thread(vA, nA, G.lut, lutA, mk(-v.x, -v.y, -v.z), c_dots);
            const V3 dB = mk(fmaf(r0, v.x, fmaf(r3, v.y, r6 * v.z)), fmaf(r1, v.x, fmaf(r4, v.y, r7 * v.z)),
                             fmaf(r2, v.x, fmaf(r5, v.y, r8 * v.z)));  // R^T v
            const int sb = support_thread(vB, nB, G.lut, lutB, dB, c_dots);
            if (COUNT) c_iters++;
            const float4 pa = vA[sa], pb = vB[sb];
            const V3 w = mk(pa.x - fmaf(r0, pb.x, fmaf(r1, pb.y, fmaf(r2, pb.z, tx))),
                            pa.y - fmaf(r3, pb.x, fmaf(r4, pb.y, fmaf(r5, pb.z, ty))),
                            pa.z - fmaf(r6, pb.x, fmaf(r7, pb.y, fmaf(r8, pb.z, tz))));
            const int id = (sa << 16) | sb;
            if (!have_point) {  // the first iteration seeds the simplex with a real point of A - B
#if GJK_FIRST_PRUNE
                // w minimises v . x over A - B for ANY direction v, so the separating-plane bound already holds for the
                // line of centres: most pairs the planning could not decide with its table-based widths end here, after
                // one support query instead of two (the lane reports nothing: its upper bound is above the limit too)
                const float vw0 = dot(v, w);
                if (vw0 > 0.0f && vw0 * vw0 >= lim * lim * vv * (1.0f + 4e-6f)) {
                    done = true;
                } else
#endif
                {
                    S.p0 = w; S.i0 = id; S.n = 1;
                    v = w; vv = dot(v, v);
                    have_point = true;
                    if (touch >= 0.0f && vv <= touch * touch) done = true;
                    if (vv <= 1e-20f) { vv = 0.0f; done = true; }
                }
            } else {
                const float vw = dot(v, w);
                const float nv = sqrtf(vv);
                if (vw > 0.0f && vw * vw >= lim * lim * vv) done = true;                      // pruned
                else if (vv - vw <= fmaxf(1e-6f * vv, 3e-7f * nv)) done = true;               // converged
                else if (id == S.i0 || id == S.i1 || id == S.i2 || id == S.i3) done = true;   // support repeats
                else {
                    if (S.n == 1) { S.p1 = w; S.i1 = id; }
                    else if (S.n == 2) { S.p2 = w; S.i2 = id; }
                    else { S.p3 = w; S.i3 = id; }
                    S.n++;
                    V3 nvv;
                    if (simplex_solve(S, nvv)) { vv = 0.0f; done = true; }
                    else {
                        if (S.n < 4) S.i3 = -1;
                        if (S.n < 3) S.i2 = -1;
                        if (S.n < 2) S.i1 = -1;
                        const float nd = dot(nvv, nvv);
                        if (!(nd < vv)) done = true;  // no progress (numerical floor) or a NaN
                        else {
                            v = nvv; vv = nd;
                            if (vv <= 1e-20f) { vv = 0.0f; done = true; }
                            if (touch >= 0.0f && vv <= touch * touch) done = true;
                        }
                    }
                }
            }
            if (++it >= 32) done = true;
        }
        // ---------------- the group's best upper bound tightens every member's pruning limit
        const float d = sqrtf(vv) - m;
        const unsigned gb = __reduce_min_sync(peers, (dist && have_point) ? fkey(d) : 0xffffffffu);
        if (dist && gb != 0xffffffffu) lim = fminf(lim, funkey(gb) + m);
        // ---------------- a finished pair reports its last upper bound (the exact distance if it converged)
        if (done) {
            if (have_point && d <= thr) {
                if (cls >= GJK_CONTACT) atomicMin(&A.res[(size_t)env * SM_RES_STRIDE + cls], meta >> 8);
                else if (d + m <= lim) atomicMin(&A.res[(size_t)env * SM_RES_STRIDE + cls], fkey(d));  // group's best
            }
            busy = false;
            if (COUNT && A.counters)   // histogram of the iterations per pair: <= 4, 8, 12, 16, 24, more
                atomicAdd(&A.counters[10 + (it <= 4 ? 0 : it <= 8 ? 1 : it <= 12 ? 2 : it <= 16 ? 3 : it <= 24 ? 4 : 5)], 1ull);
        }
    }
    if (COUNT && A.counters) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c_calls += __shfl_xor_sync(FULL, c_calls, o);
            c_iters += __shfl_xor_sync(FULL, c_iters, o);
            c_dots += __shfl_xor_sync(FULL, c_dots, o);
        }
        if (lane == 0) {
            atomicAdd(&A.counters[0], (unsigned long long)c_calls);
            atomicAdd(&A.counters[1], (unsigned long long)c_iters);
            atomicAdd(&A.counters[2], (unsigned long long)c_dots);
        }
    }
}
