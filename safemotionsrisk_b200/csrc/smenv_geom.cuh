// smenv_geom.cuh -- float32 geometry of the env step: rigid transforms, forward kinematics, warp-cooperative GJK.
//
// Everything here is written for a small instruction footprint (each helper exists once, __noinline__): the first
// fused version of the step kernel was 350 KB of SASS and spent ~48 cycles per issue waiting for instruction fetch.
#pragma once
#include "smenv_device.cuh"

struct Xf {
    float r[9];
    float t[3];
};
struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ V3 xf_apply(const Xf& T, float x, float y, float z) {
    return mk(fmaf(T.r[0], x, fmaf(T.r[1], y, fmaf(T.r[2], z, T.t[0]))),
              fmaf(T.r[3], x, fmaf(T.r[4], y, fmaf(T.r[5], z, T.t[1]))),
              fmaf(T.r[6], x, fmaf(T.r[7], y, fmaf(T.r[8], z, T.t[2]))));
}
__device__ __forceinline__ V3 xf_rot_t(const Xf& T, V3 d) {  // R^T d
    return mk(fmaf(T.r[0], d.x, fmaf(T.r[3], d.y, T.r[6] * d.z)), fmaf(T.r[1], d.x, fmaf(T.r[4], d.y, T.r[7] * d.z)),
              fmaf(T.r[2], d.x, fmaf(T.r[5], d.y, T.r[8] * d.z)));
}
// C = A o B  (apply B first)
__device__ __forceinline__ void xf_compose(const Xf& A, const Xf& B, Xf& C) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
            C.r[3 * i + k] = fmaf(A.r[3 * i], B.r[k], fmaf(A.r[3 * i + 1], B.r[3 + k], A.r[3 * i + 2] * B.r[6 + k]));
        C.t[i] = fmaf(A.r[3 * i], B.t[0], fmaf(A.r[3 * i + 1], B.t[1], fmaf(A.r[3 * i + 2], B.t[2], A.t[i])));
    }
}
__device__ __forceinline__ void xf_identity(Xf& T) {
#pragma unroll
    for (int i = 0; i < 9; ++i) T.r[i] = (i % 4 == 0) ? 1.0f : 0.0f;
    T.t[0] = T.t[1] = T.t[2] = 0.0f;
}
__device__ __forceinline__ void axis_angle(float x, float y, float z, float c, float s, float* R) {
    float t = 1.0f - c;
    R[0] = t * x * x + c;     R[1] = t * x * y - s * z; R[2] = t * x * z + s * y;
    R[3] = t * x * y + s * z; R[4] = t * y * y + c;     R[5] = t * y * z - s * x;
    R[6] = t * x * z - s * y; R[7] = t * y * z + s * x; R[8] = t * z * z + c;
}
__device__ __forceinline__ void quat_to_mat(float4 q, float* R) {
    float x = q.x, y = q.y, z = q.z, w = q.w;
    float n = x * x + y * y + z * z + w * w;
    float s = n > 0.0f ? 2.0f / n : 0.0f;
    R[0] = 1.0f - s * (y * y + z * z); R[1] = s * (x * y - w * z);        R[2] = s * (x * z + w * y);
    R[3] = s * (x * y + w * z);        R[4] = 1.0f - s * (x * x + z * z); R[5] = s * (y * z - w * x);
    R[6] = s * (x * z - w * y);        R[7] = s * (y * z + w * x);        R[8] = 1.0f - s * (x * x + y * y);
}
__device__ __noinline__ void euler_to_mat(float e0, float e1, float e2, float* R) {
    float sr, cr, sp, cp, sy, cy;
    sincosf(e0, &sr, &cr); sincosf(e1, &sp, &cp); sincosf(e2, &sy, &cy);
    R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
    R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
    R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
}

// ------------------------------------------------------------------------------------------------------------------
// shared-memory copy of the scene tables that lanes index with different values (constant memory would serialise)
// ------------------------------------------------------------------------------------------------------------------
struct SceneSmem {
    DevShape shapes[SM_MAX_SHAPES];
    float jR[SM_MAX_JOINTS][9], jt[SM_MAX_JOINTS][3], jaxis[SM_MAX_JOINTS][3];
    short static_pairs[SM_MAX_PAIRS][2], self_pairs[SM_MAX_PAIRS][2];
    short mov_reward[SM_MAX_MOV_ROBOT], mov_contact[SM_MAX_MOV_ROBOT];
    float contact_thresh[SM_MAX_OBSTACLES][SM_MAX_MOV_ROBOT];
};

// What a geometry CTA stages into shared memory, as one contiguous image in global memory (built on the host in
// smenv_create): 16-byte vector copies instead of divergent (serialised) constant-memory reads.
struct SceneImage {
    SceneSmem scene;
    uint32_t pair_tab[SM_MAX_PLAN_PAIRS];   // shape A | shape B << 12 | class << 24 | (B is a static box in the world frame) << 31
    float2 pair_rm[SM_MAX_PLAN_PAIRS];      // x: sum of the margins (rounded down), y: what the sphere bound subtracts
                                            // (radii + margins; static box: radius of A + margins), rounded up
};
static_assert(sizeof(SceneImage) % 16 == 0, "SceneImage is copied in 16-byte vectors");

// ------------------------------------------------------------------------------------------------------------------
// forward kinematics of the serial chain by a warp scan:  lane j builds L_j = [R_fix Rot(axis, q_j) | t_fix], an
// inclusive scan under composition over lanes 0..7 (three shuffle rounds) yields T_j = L_0 o ... o L_j, the frame of
// the link driven by joint j (ctlp.py:2940-2988; LinkBase.get_position :5163-5195 -> getLinkState[4:6]).
// out[0] = world, out[1 + j] = T_j.  c, s: cos / sin of the lane's own joint angle.
// ------------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void fk_scan(const SceneSmem& sm, float c, float s, Xf* out, int lane) {
    const int j = lane & 7;
    Xf X;
    {
        float Rj[9];
        axis_angle(sm.jaxis[j][0], sm.jaxis[j][1], sm.jaxis[j][2], c, s, Rj);
        const float* A = sm.jR[j];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k)
                X.r[3 * i + k] = fmaf(A[3 * i], Rj[k], fmaf(A[3 * i + 1], Rj[3 + k], A[3 * i + 2] * Rj[6 + k]));
        X.t[0] = sm.jt[j][0]; X.t[1] = sm.jt[j][1]; X.t[2] = sm.jt[j][2];
    }
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
        Xf P, C;
#pragma unroll
        for (int i = 0; i < 9; ++i) P.r[i] = __shfl_up_sync(FULL, X.r[i], d, 8);
#pragma unroll
        for (int i = 0; i < 3; ++i) P.t[i] = __shfl_up_sync(FULL, X.t[i], d, 8);
        xf_compose(P, X, C);
        if (j >= d) X = C;
    }
    if (lane == 0) xf_identity(out[0]);
    if (lane < c_sc.n_joints) out[1 + lane] = X;
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------------------------
// warp-cooperative GJK distance between two convex vertex sets held in shared memory
// (restates what p.getClosestPoints computes on the margin-less cores; call sites ctlp.py:3267, :3300, :3353)
// ------------------------------------------------------------------------------------------------------------------
struct GjkCounters {
    unsigned calls, iters, dots;
    float* trace;  // debug: 8 floats per iteration (smenv_debug_gjk), else NULL
};

__device__ __forceinline__ unsigned fkey(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// support vertex index of a hull in local direction d; the 32 lanes split the vertices
__device__ __forceinline__ int warp_support(const float4* __restrict__ v, int n, V3 d, int lane) {
    float best = -FLT_MAX;
    int bi = 0;
    // hulls have at most 255 vertices: eight independent loads per lane, issued together (the vertices of the planning
    // kernels come from global memory / L2: one latency instead of eight)
#pragma unroll 8
    for (int i = lane; i < n; i += 32) {
        float4 p = v[i];
        float s = fmaf(p.x, d.x, fmaf(p.y, d.y, p.z * d.z));
        if (s > best) { best = s; bi = i; }
    }
    unsigned key = fkey(best);
    unsigned mx = __reduce_max_sync(FULL, key);
    unsigned bal = __ballot_sync(FULL, key == mx);
    return __shfl_sync(FULL, bi, __ffs(bal) - 1);
}

// closest point to the origin on segment ab; w = vertex mask (bit 0 = a kept, bit 1 = b kept)
__device__ __forceinline__ float4 closest_segment(V3 a, V3 b) {
    V3 ab = b - a;
    float t = -dot(a, ab), den = dot(ab, ab);
    if (t <= 0.0f || !(den > 0.0f)) return make_float4(a.x, a.y, a.z, __int_as_float(1));
    if (t >= den) return make_float4(b.x, b.y, b.z, __int_as_float(2));
    V3 p = a + (t / den) * ab;
    return make_float4(p.x, p.y, p.z, __int_as_float(3));
}

struct Simplex {
    V3 p0, p1, p2, p3;
    int i0, i1, i2, i3;  // (vertex of A << 16 | vertex of B) of each simplex point
    int n;
};

// Candidates of the closest boundary point of a simplex to the origin.  Robust float32 formulation: the answer is the
// foot of a perpendicular onto a face (when it falls inside the face) or the closest point of an edge; ALL candidates
// are evaluated and the nearest wins.  Voronoi-region tests (Ericson 5.1.5) misclassify thin triangles in float32 and
// then return a point that is not the minimum, which stalls GJK with a support vertex that is "already in the
// simplex" (seen as 1.6e-4 m distance errors on the table).
//   bits: the simplex vertices the candidate keeps (bit i = point i)
__device__ __forceinline__ void edge_candidate(V3 a, V3 b, int bit_a, int bit_b, float& qb, V3& bv, int& bm) {
    const float4 p = closest_segment(a, b);
    const float q = p.x * p.x + p.y * p.y + p.z * p.z;
    if (q < qb) {
        const int m = __float_as_int(p.w);
        qb = q; bv = mk(p.x, p.y, p.z);
        bm = ((m & 1) ? bit_a : 0) | ((m & 2) ? bit_b : 0);
    }
}
__device__ __forceinline__ void face_candidate(V3 a, V3 b, V3 c, int bits, float& qb, V3& bv, int& bm) {
    const V3 ab = b - a, ac = c - a;
    const V3 n = cross(ab, ac);
    const float nn = dot(n, n);
    if (nn > 1e-12f * dot(ab, ab) * dot(ac, ac)) {  // usable area: foot of the perpendicular p = n (a.n) / |n|^2
        const float s = dot(a, n) / nn;
        const V3 p = s * n;
        // inside test with the edge functions, all against the same normal
        const float w0 = dot(cross(b - p, c - p), n), w1 = dot(cross(c - p, a - p), n), w2 = dot(cross(a - p, b - p), n);
        const float qi = dot(p, p);
        if (w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f && qi < qb) { qb = qi; bv = p; bm = bits; }
    }
}
// does the origin lie on the same side of face abc as the fourth vertex d?
__device__ __forceinline__ bool same_side(V3 a, V3 b, V3 c, V3 d) {
    const V3 nrm = cross(b - a, c - a);
    const float sd = dot(d - a, nrm), so = -dot(a, nrm);
    return so * sd > 0.0f;
}

// closest point of the simplex to the origin; reduces the simplex to the supporting face.  true = origin enclosed.
// One straight-line body for the segment, triangle and tetrahedron cases (6 edges, 4 faces, each evaluated once): the
// lanes of a warp hold simplices of different sizes, and a body per case would run the three one after the other
// (the tetrahedron, done face by face, evaluated every edge twice).
__device__ __forceinline__ bool simplex_solve(Simplex& S, V3& v) {
    const V3 A = S.p0, B = S.p1, C = S.p2, D = S.p3;
    const int ia = S.i0, ib = S.i1, ic = S.i2, id = S.i3;
    const int n = S.n;
    float qb = FLT_MAX;
    V3 bv = mk(0.f, 0.f, 0.f);
    int bm = 1;
    edge_candidate(A, B, 1, 2, qb, bv, bm);
    if (n >= 3) {
        edge_candidate(A, C, 1, 4, qb, bv, bm);
        edge_candidate(B, C, 2, 4, qb, bv, bm);
        if (n >= 4) {
            edge_candidate(A, D, 1, 8, qb, bv, bm);
            edge_candidate(B, D, 2, 8, qb, bv, bm);
            edge_candidate(C, D, 4, 8, qb, bv, bm);
        }
        face_candidate(A, B, C, 7, qb, bv, bm);
        if (n >= 4) {
            face_candidate(A, B, D, 11, qb, bv, bm);
            face_candidate(A, C, D, 13, qb, bv, bm);
            face_candidate(B, C, D, 14, qb, bv, bm);
            // The origin counts as enclosed only if every face test says "inside" AND the tetrahedron is not flat: in
            // float32 a sliver of four nearly coplanar support points must never certify a penetration.
            if (same_side(A, B, C, D) && same_side(A, B, D, C) && same_side(A, C, D, B) && same_side(B, C, D, A)) {
                const V3 e1 = B - A, e2 = C - A, e3 = D - A;
                const float det = dot(e3, cross(e1, e2));
                const float scale2 = dot(e1, e1) * dot(e2, e2) * dot(e3, e3);
                if (det * det > 1e-8f * scale2) return true;  // normalised volume above 1e-4: a genuine enclosure
            }
        }
    }
    // keep the vertices of the supporting feature, in their order
    int k = 0;
    if (bm & 1) k = 1;
    if (bm & 2) { if (k == 0) { S.p0 = B; S.i0 = ib; } else { S.p1 = B; S.i1 = ib; } ++k; }
    if (bm & 4) {
        if (k == 0) { S.p0 = C; S.i0 = ic; } else if (k == 1) { S.p1 = C; S.i1 = ic; } else { S.p2 = C; S.i2 = ic; }
        ++k;
    }
    if (bm & 8) {
        if (k == 0) { S.p0 = D; S.i0 = id; } else if (k == 1) { S.p1 = D; S.i1 = id; } else { S.p2 = D; S.i2 = id; }
        ++k;
    }
    (void)ia;
    S.n = k;
    v = bv;
    return false;
}

__device__ __forceinline__ const Xf* frame_ptr(const DevShape& sh, const Xf* robot, const Xf* obst) {
    return sh.frame >= 100 ? obst + (sh.frame - 100) : robot + sh.frame;
}

// Distance between shapes ia and ib MINUS both collision margins (Bullet's getClosestPoints metric).
//   upper  > 0: stop as soon as the distance is proven >= upper (returns a value >= upper): exact pruning of pairs
//               that cannot lower the running minimum / cannot be inside the query distance.
//   touch >= 0: stop as soon as the distance is proven <= touch (returns a value <= touch): contact tests.
// All 32 lanes call together; frames come from shared memory.
__device__ __noinline__ float pair_distance(const float4* __restrict__ verts, const SceneSmem& sm, int ia, int ib,
                                            const Xf* robot, const Xf* obst, float upper, float touch, int lane,
                                            GjkCounters* cnt) {
    const DevShape& A = sm.shapes[ia];
    const DevShape& B = sm.shapes[ib];
    const Xf TA = *frame_ptr(A, robot, obst);
    const Xf TB = *frame_ptr(B, robot, obst);
    const float4* vA = verts + A.off;
    const float4* vB = verts + B.off;
    const int nA = A.cnt, nB = B.cnt;
    const float m = A.margin + B.margin;
    if (upper > 0.0f) upper += m;
    if (touch >= 0.0f) touch += m;
    Simplex S;
    S.n = 0;
    S.i0 = S.i1 = S.i2 = S.i3 = -1;
    S.p0 = S.p1 = S.p2 = S.p3 = mk(0.f, 0.f, 0.f);
    V3 v = xf_apply(TA, A.cx, A.cy, A.cz) - xf_apply(TB, B.cx, B.cy, B.cz);  // first direction: between the centres
    float vv = dot(v, v);
    if (vv < 1e-12f) { v = mk(1.f, 0.f, 0.f); vv = 1.f; }
    bool have_point = false;
    if (cnt) cnt->calls++;
#pragma unroll 1
    for (int it = 0; it < 32; ++it) {
        V3 dA = xf_rot_t(TA, mk(-v.x, -v.y, -v.z));
        V3 dB = xf_rot_t(TB, v);
        int sa = warp_support(vA, nA, dA, lane);
        int sb = warp_support(vB, nB, dB, lane);
        if (cnt) { cnt->iters++; cnt->dots += (unsigned)(nA + nB); }
        float4 pa = vA[sa], pb = vB[sb];
        V3 w = xf_apply(TA, pa.x, pa.y, pa.z) - xf_apply(TB, pb.x, pb.y, pb.z);
        int id = (sa << 16) | sb;
        if (!have_point) {  // the first iteration only seeds the simplex with a real point of A - B
            S.p0 = w; S.i0 = id; S.n = 1;
            v = w; vv = dot(v, v);
            have_point = true;
            if (touch >= 0.0f && vv <= touch * touch) break;
            if (vv <= 1e-20f) { vv = 0.0f; break; }
            continue;
        }
        float vw = dot(v, w);
        if (cnt && cnt->trace && lane == 0) {
            float* tr = cnt->trace + 8 * it;
            tr[0] = (float)S.n; tr[1] = vv; tr[2] = vw; tr[3] = (float)sa; tr[4] = (float)sb; tr[5] = v.x; tr[6] = v.y; tr[7] = v.z;
        }
        if (upper > 0.0f && vw > 0.0f && vw * vw >= upper * upper * vv) { vv = fmaxf(vv, upper * upper); break; }
        float nv = sqrtf(vv);
        if (vv - vw <= fmaxf(1e-6f * vv, 3e-7f * nv)) break;                     // converged
        if (id == S.i0 || id == S.i1 || id == S.i2 || id == S.i3) break;         // support already in the simplex
        if (S.n == 1) { S.p1 = w; S.i1 = id; }
        else if (S.n == 2) { S.p2 = w; S.i2 = id; }
        else { S.p3 = w; S.i3 = id; }
        S.n++;
        V3 nvv;
        if (simplex_solve(S, nvv)) { vv = 0.0f; break; }
        if (S.n < 4) S.i3 = -1;
        if (S.n < 3) S.i2 = -1;
        if (S.n < 2) S.i1 = -1;
        float nd = dot(nvv, nvv);
        if (!(nd < vv)) break;  // no progress (numerical floor) or a NaN from a degenerate simplex
        v = nvv; vv = nd;
        if (vv <= 1e-20f) { vv = 0.0f; break; }
        if (touch >= 0.0f && vv <= touch * touch) break;
    }
    return sqrtf(vv) - m;
}

// bounding-sphere lower bound of the same quantity (per lane, any pair)
__device__ __forceinline__ float pair_lower_bound(const SceneSmem& sm, int ia, int ib, const Xf* robot, const Xf* obst) {
    const DevShape& A = sm.shapes[ia];
    const DevShape& B = sm.shapes[ib];
    const Xf* TA = frame_ptr(A, robot, obst);
    V3 ca = xf_apply(*TA, A.cx, A.cy, A.cz);
    if (B.frame == 0) {  // static shape in the world frame: sphere against its axis-aligned box (tight for the table)
        float dx = fmaxf(fmaxf(B.bmin[0] - ca.x, ca.x - B.bmax[0]), 0.f);
        float dy = fmaxf(fmaxf(B.bmin[1] - ca.y, ca.y - B.bmax[1]), 0.f);
        float dz = fmaxf(fmaxf(B.bmin[2] - ca.z, ca.z - B.bmax[2]), 0.f);
        return sqrtf(dx * dx + dy * dy + dz * dz) - A.radius - A.margin - B.margin;
    }
    const Xf* TB = frame_ptr(B, robot, obst);
    V3 d = ca - xf_apply(*TB, B.cx, B.cy, B.cz);
    return sqrtf(dot(d, d)) - A.radius - B.radius - A.margin - B.margin;
}
