// smenv_plan.cuh -- planning kernels of the env step: bounding-volume culling that turns each env's collision queries
// into GJK work items (smenv_gjk.cuh).  One warp per env, no hull vertices touched here.
//
//   contact_plan_kernel   lane k = physics sub-step k+1: serial FK of the motor-tracked pose, robot link spheres
//                         against the obstacle spheres (whole body first, then per convex part), one contact item per
//                         (sub-step, link shape, obstacle part) whose spheres come within the manifold threshold
//                         (ObstacleWrapperSim.update, ctlp.py:2590-2862; contact semantics SURVEY Appendix B.5)
//   distance_plan_kernel  obstacle poses at the end of the step, FK of the new knot (warp scan), then lanes = pairs:
//                         sphere lower bound and centroid upper bound per pair, class-wise best upper bound, one
//                         distance item per pair that could still hold the class minimum
//                         (get_minimum_distance ctlp.py:3282-3374, .._to_moving_obstacles :3217-3280)
#pragma once
#include "smenv_gjk.cuh"
#include "smenv_joint.cuh"

struct PlanArgs {
    SmBuffers buf;
    int n;
    const float* scratch;    // [n][SM_SCRATCH_FLOATS] sub-step poses written by the joint kernels
    GjkItem* items;
    int* item_count;         // device counter
    int capacity;
    int* overflow;           // device flag: items that did not fit (the step's results are then invalid)
    unsigned* res;           // [n][SM_RES_STRIDE]
    const double* kin;       // kinematic records to plan for (buf.kin in the step, the caller's array in the hook)
    const double* obst;      // obstacle records
    int advance;             // 1: obstacle poses at the end of the step about to be finished; 0: poses as recorded
    int* cwork;              // [0] = count, [1..] = spans (env * 8 + span) whose contacts need the fine planning
    double* target;          // [n][SM_TP_STRIDE] target-point records (the step only), or NULL
    const double* hkin;      // Human scene: [n][SM_KIN_STRIDE] joint state of the human (its knot pose is the obstacle pose)
    unsigned long long* counters;
};

// sub-steps (1-based) of a ball that still test contacts: all before the counters retire it (ctlp.py:2840-2848)
__device__ __forceinline__ int ball_k_end(const double* ob) {
    const int idx0 = (int)ob[SM_OB_INDEX];
    int k1 = (int)ob[SM_OB_BALL_NMAX] - idx0 + 1, k2 = (int)ob[SM_OB_BALL_NHIT] - idx0;
    int k_end = k1 < k2 ? k1 : k2;
    return k_end < 1 ? 1 : k_end;
}

// pose of B in the frame of A
__device__ __forceinline__ void rel_pose(const Xf& TA, const Xf& TB, float* R, float* t) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
            R[3 * i + k] = fmaf(TA.r[i], TB.r[k], fmaf(TA.r[3 + i], TB.r[3 + k], TA.r[6 + i] * TB.r[6 + k]));
        const float dx = TB.t[0] - TA.t[0], dy = TB.t[1] - TA.t[1], dz = TB.t[2] - TA.t[2];
        t[i] = fmaf(TA.r[i], dx, fmaf(TA.r[3 + i], dy, TA.r[6 + i] * dz));
    }
}

// squared distance of point p to the axis-aligned box [lo, hi]
__device__ __forceinline__ float box_dist2(V3 p, const float* lo, const float* hi) {
    const float dx = fmaxf(fmaxf(lo[0] - p.x, p.x - hi[0]), 0.f);
    const float dy = fmaxf(fmaxf(lo[1] - p.y, p.y - hi[1]), 0.f);
    const float dz = fmaxf(fmaxf(lo[2] - p.z, p.z - hi[2]), 0.f);
    return fmaf(dx, dx, fmaf(dy, dy, dz * dz));
}

// Lower bound of the distance between shapes A and B (minus both margins) from the support-width tables: separation
// along the line between the bounding-sphere centres.
__device__ __forceinline__ float axis_lower_bound_d(const DevShape& SA, const DevShape& SB, const Xf& TA, const Xf& TB,
                                                    V3 d /* centre of B - centre of A, world frame */) {
    const float dist = sqrtf(dot(d, d));
    if (!(dist > 1e-9f)) return -FLT_MAX;
    const V3 da = xf_rot_t(TA, d), db = xf_rot_t(TB, mk(-d.x, -d.y, -d.z));
    const float ha = __ldg(c_sc.hwidth + SA.hw + lut_cell(da.x, da.y, da.z));
    const float hb = __ldg(c_sc.hwidth + SB.hw + lut_cell(db.x, db.y, db.z));
    return dist * (1.0f - 1e-6f) - ha - hb - SA.margin - SB.margin;
}
__device__ __forceinline__ float axis_lower_bound(const DevShape& SA, const DevShape& SB, const Xf& TA, const Xf& TB) {
    return axis_lower_bound_d(SA, SB, TA, TB, xf_apply(TB, SB.cx, SB.cy, SB.cz) - xf_apply(TA, SA.cx, SA.cy, SA.cz));
}

__device__ __forceinline__ void write_item(GjkItem* dst, int env, int ia, int ib, int cls, int sub, float thr,
                                           const Xf& TA, const Xf& TB) {
    float R[9], t[3];
    rel_pose(TA, TB, R, t);
    float4* o = reinterpret_cast<float4*>(dst);
    o[0] = make_float4(__int_as_float(env), __uint_as_float((unsigned)ia | ((unsigned)ib << 16)),
                       __uint_as_float((unsigned)cls | ((unsigned)sub << 8)), thr);
    o[1] = make_float4(R[0], R[1], R[2], R[3]);
    o[2] = make_float4(R[4], R[5], R[6], R[7]);
    o[3] = make_float4(R[8], t[0], t[1], t[2]);
}

__device__ __forceinline__ int warp_exclusive_sum(int x, int lane, int& total) {
    int s = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(FULL, s, o);
        if (lane >= o) s += y;
    }
    total = __shfl_sync(FULL, s, 31);
    return s - x;
}

// one link of the serial chain: F <- F o [R_fix Rot(axis_j, q) | t_fix]
__device__ __forceinline__ void fk_chain_step(const SceneSmem& sm, Xf& F, int j, float q) {
    float s, c;
    sincosf(q, &s, &c);
    Xf L, C;
    float Rj[9];
    axis_angle(sm.jaxis[j][0], sm.jaxis[j][1], sm.jaxis[j][2], c, s, Rj);
    if ((c_sc.jr_identity >> j) & 1) {   // uniform: no fixed rotation in front of the joint (x * 1 + 0 drops out exactly)
#pragma unroll
        for (int i = 0; i < 9; ++i) L.r[i] = Rj[i];
    } else {
        const float* A = sm.jR[j];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k)
                L.r[3 * i + k] = fmaf(A[3 * i], Rj[k], fmaf(A[3 * i + 1], Rj[3 + k], A[3 * i + 2] * Rj[6 + k]));
    }
    L.t[0] = sm.jt[j][0]; L.t[1] = sm.jt[j][1]; L.t[2] = sm.jt[j][2];
    xf_compose(F, L, C);
    F = C;
}

// ------------------------------------------------------------------------------------------------------------------
// Coarse contact phase: 8 threads per env, each clears a span of ceil(S / 8) consecutive sub-steps with ONE forward
// kinematics (the pose of the span's middle sub-step): every contact sphere is inflated by what the joints move inside
// the span (sum_j |dq_j| * contact_rho[slot][j]) and tested against the obstacle's whole-body sphere at each sub-step
// of the span.  Only envs with a span that cannot be cleared go to the fine planning (contact_plan_kernel), which
// needs a full warp per env; typically that is a small fraction of the envs.
// ------------------------------------------------------------------------------------------------------------------
#define SM_COARSE_LANES 8
#define SM_COARSE_SPAN 4 /* max sub-steps per coarse thread: ceil(SM_MAX_SUB / SM_COARSE_LANES) */

template <bool COUNT>
__global__ void __launch_bounds__(256) contact_coarse_kernel(PlanArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemLayout L = block_prologue(smem_raw, false);
    const SceneSmem& sm = L.bs->scene;
    const int tid = threadIdx.x, lane = tid & 31;
    const int t = blockIdx.x * blockDim.x + tid;
    const int env = t / SM_COARSE_LANES, c = t % SM_COARSE_LANES;
    const int S = c_sc.substeps, stride = c_sc.contact_stride;
    const int kind = c_sc.n_obstacles > 0 ? c_sc.obst_kind[0] : SM_OBST_NONE;
    const int span = (S + SM_COARSE_LANES - 1) / SM_COARSE_LANES;
    bool flag = false;
    if (env < A.n && stride > 0 && c * span < S) {
        const double* ob = A.obst + (size_t)env * SM_OBST_STRIDE;
        const int k0 = c * span, k1 = k0 + span < S ? k0 + span : S;
        const int kc = (k0 + k1 - 1) >> 1;
        V3 oc[2][SM_COARSE_SPAN];
        unsigned valid = 0;  // bit (o * SM_COARSE_SPAN + i): obstacle o is tested at sub-step k0 + i
        if (ob[SM_OB_LATCH] == 0.0) {
            const int idx0 = (int)ob[SM_OB_INDEX];
            if (kind == SM_OBST_PLANET && c_sc.terminate_moving) {
#pragma unroll
                for (int i = 0; i < SM_COARSE_SPAN; ++i) {
                    if (k0 + i >= k1) continue;
                    const int idx = (idx0 + k0 + i) % c_sc.planet_steps;
                    const float4 p0 = __ldg(c_sc.planet_pos[0] + idx);
                    oc[0][i] = mk(p0.x, p0.y, p0.z);
                    valid |= 1u << i;
                    if (c_sc.n_obstacles > 1) {
                        int id1 = (idx + c_sc.planet_shift) % c_sc.planet_steps;
                        if (id1 < 0) id1 += c_sc.planet_steps;
                        const float4 p1 = __ldg(c_sc.planet_pos[1] + id1);
                        oc[1][i] = mk(p1.x, p1.y, p1.z);
                        valid |= 1u << (SM_COARSE_SPAN + i);
                    }
                }
            } else if (kind == SM_OBST_BALL && ob[SM_OB_BALL_ACTIVE] != 0.0) {
                const double dt = xdiv(c_sc.ts, (double)S);
                const int k_end = ball_k_end(ob);
                const double ball_t = ob[SM_OB_BALL_T];
#pragma unroll
                for (int i = 0; i < SM_COARSE_SPAN; ++i) {
                    const int k = k0 + i, sub = k + 1;
                    if (k >= k1 || sub > k_end - 1) continue;
                    const double tn = ball_t + (double)sub * dt;  // active-area test on the new position
                    const double px = ob[SM_OB_BALL_P0] + ob[SM_OB_BALL_V0] * tn;
                    const double py = ob[SM_OB_BALL_P0 + 1] + ob[SM_OB_BALL_V0 + 1] * tn;
                    if (!(sqrt(px * px + py * py) < c_sc.ball_active_xy)) continue;
                    const double tk = ball_t + (double)k * dt;    // manifold of the previous position
                    oc[0][i] = mk((float)(ob[SM_OB_BALL_P0] + ob[SM_OB_BALL_V0] * tk),
                                  (float)(ob[SM_OB_BALL_P0 + 1] + ob[SM_OB_BALL_V0 + 1] * tk),
                                  (float)(ob[SM_OB_BALL_P0 + 2] + ob[SM_OB_BALL_V0 + 2] * tk + (0.5 * -9.81) * (tk * tk)));
                    valid |= 1u << i;
                }
            }
        }
        // one sphere per obstacle for the whole span: centre = mean of the tested sub-step positions, radius grown by
        // the largest deviation from it (a planet travels ~1 cm in a span, the ball ~7 cm)
        V3 om[2];
        float otr[2] = {0.f, 0.f};
#pragma unroll
        for (int o = 0; o < 2; ++o) {
            V3 sum = mk(0.f, 0.f, 0.f);
            int cnt = 0;
#pragma unroll
            for (int i = 0; i < SM_COARSE_SPAN; ++i)
                if ((valid >> (o * SM_COARSE_SPAN + i)) & 1u) { sum = sum + oc[o][i]; ++cnt; }
            om[o] = cnt ? (1.0f / (float)cnt) * sum : mk(0.f, 0.f, 0.f);
            float d2 = 0.f;
#pragma unroll
            for (int i = 0; i < SM_COARSE_SPAN; ++i)
                if ((valid >> (o * SM_COARSE_SPAN + i)) & 1u) { const V3 e = oc[o][i] - om[o]; d2 = fmaxf(d2, dot(e, e)); }
            otr[o] = sqrtf(d2) * (1.0f + 1e-6f) + 1e-6f;
        }
        const unsigned omask = ((valid & ((1u << SM_COARSE_SPAN) - 1u)) ? 1u : 0u) | ((valid >> SM_COARSE_SPAN) ? 2u : 0u);
        if (valid) {
            const float* scr = A.scratch + (size_t)env * SM_SCRATCH_FLOATS;
            float dq[SM_MAX_JOINTS];
            float qc[SM_MAX_JOINTS];
            static_assert(SM_MAX_JOINTS == 8, "rows of the sub-step poses are read as two float4");
            {   // rows of 8 floats, 32-byte aligned: two 16-byte loads per sub-step
                const float4* rows = reinterpret_cast<const float4*>(scr);
                const float4 c0 = rows[2 * kc], c1 = rows[2 * kc + 1];
                qc[0] = c0.x; qc[1] = c0.y; qc[2] = c0.z; qc[3] = c0.w; qc[4] = c1.x; qc[5] = c1.y; qc[6] = c1.z; qc[7] = c1.w;
#pragma unroll
                for (int j = 0; j < SM_MAX_JOINTS; ++j) dq[j] = 0.f;
#pragma unroll
                for (int i = 0; i < SM_COARSE_SPAN; ++i) {
                    if (k0 + i >= k1) continue;
                    const float4 a0 = rows[2 * (k0 + i)], a1 = rows[2 * (k0 + i) + 1];
                    dq[0] = fmaxf(dq[0], fabsf(a0.x - qc[0])); dq[1] = fmaxf(dq[1], fabsf(a0.y - qc[1]));
                    dq[2] = fmaxf(dq[2], fabsf(a0.z - qc[2])); dq[3] = fmaxf(dq[3], fabsf(a0.w - qc[3]));
                    dq[4] = fmaxf(dq[4], fabsf(a1.x - qc[4])); dq[5] = fmaxf(dq[5], fabsf(a1.y - qc[5]));
                    dq[6] = fmaxf(dq[6], fabsf(a1.z - qc[6])); dq[7] = fmaxf(dq[7], fabsf(a1.w - qc[7]));
                }
            }
            Xf F;
            xf_identity(F);
#pragma unroll 1
            for (int f = 0; f <= c_sc.n_joints && !flag; ++f) {
                if (f > 0) fk_chain_step(sm, F, f - 1, scr[kc * SM_MAX_JOINTS + f - 1]);   // (cached) re-read: no dynamic index into registers
#pragma unroll 1
                for (int slot = c_sc.contact_frame_start[f]; slot < c_sc.contact_frame_start[f + 1]; ++slot) {
                    const DevShape& sh = sm.shapes[sm.mov_contact[slot]];
                    const V3 ctr = xf_apply(F, sh.cx, sh.cy, sh.cz);
                    float infl = 1e-5f;
#pragma unroll
                    for (int j = 0; j < SM_MAX_JOINTS; ++j) infl = fmaf(dq[j], c_sc.contact_rho[slot][j], infl);
                    const float rr = sh.radius + sh.margin + infl;
#pragma unroll
                    for (int o = 0; o < 2; ++o) {
                        if (o >= c_sc.n_obstacles) continue;  // uniform: skips the unrolled body
                        if (!((omask >> o) & 1u)) continue;
                        const float lim = rr + c_sc.obst_radius[o] + c_sc.obst_center_norm[o] + sm.contact_thresh[o][slot] + otr[o];
                        const V3 e = ctr - om[o];
                        if (dot(e, e) <= lim * lim) flag = true;
                    }
                }
            }
        }
    }
    // one list entry per span that could not be cleared: env * SM_COARSE_LANES + span
    const unsigned fm = __ballot_sync(FULL, flag);
    if (fm) {
        int base = 0;
        if (lane == __ffs(fm) - 1) base = atomicAdd(A.cwork, __popc(fm));
        base = __shfl_sync(FULL, base, __ffs(fm) - 1);
        if (flag) A.cwork[1 + base + __popc(fm & ((1u << lane) - 1u))] = t;
        if (COUNT && A.counters && lane == __ffs(fm) - 1) atomicAdd(&A.counters[5], (unsigned long long)__popc(fm));
    }
}

// ------------------------------------------------------------------------------------------------------------------
// contacts of one sub-step (one lane): serial FK chain of the tracked pose; every contact shape against the obstacle
// bounding spheres and then the spheres of the obstacle's convex parts, all inflated by the contact thresholds.
// Poses are those Bullet's collision detection of that sub-step sees: tracked robot pose before integration, obstacle
// pose of the previous update.  A pair that neither the spheres nor the separating-axis bound
// can clear becomes a contact item (appended with one atomic per item: they are rare, about one per env-step).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int contact_scan(const SceneSmem& sm, const float* __restrict__ qrow, const Xf& T0, const Xf& T1,
                                            int use_mask, const PlanArgs& A, int env, int sub) {
    int count = 0;
    const V3 oc0 = xf_apply(T0, c_sc.obst_center[0][0], c_sc.obst_center[0][1], c_sc.obst_center[0][2]);
    const V3 oc1 = xf_apply(T1, c_sc.obst_center[1][0], c_sc.obst_center[1][1], c_sc.obst_center[1][2]);
    Xf F;
    xf_identity(F);
#pragma unroll 1
    for (int f = 0; f <= c_sc.n_joints; ++f) {
        if (f > 0) fk_chain_step(sm, F, f - 1, qrow[f - 1]);  // frame f hangs off frame f-1 (checked on the host)
#pragma unroll 1
        for (int slot = c_sc.contact_frame_start[f]; slot < c_sc.contact_frame_start[f + 1]; ++slot) {
            const int ia = sm.mov_contact[slot];
            const DevShape& sh = sm.shapes[ia];
            const V3 c = xf_apply(F, sh.cx, sh.cy, sh.cz);
            const float rr = sh.radius + sh.margin;
#pragma unroll 1
            for (int o = 0; o < 2; ++o) {
                if (!((use_mask >> o) & 1)) continue;
                const float th = sm.contact_thresh[o][slot];
                const V3 dd = c - (o == 0 ? oc0 : oc1);
                const float lim = rr + c_sc.obst_radius[o] + th;
                if (dot(dd, dd) > lim * lim) continue;
                const Xf& TB = o == 0 ? T0 : T1;
                const int off = c_sc.obst_shape_off[o], cnt = c_sc.obst_shape_cnt[o];
                const V3 cl = xf_rot_t(TB, mk(c.x - TB.t[0], c.y - TB.t[1], c.z - TB.t[2]));  // in the obstacle frame
                {   // the link's sphere against the box of the whole obstacle in its own frame (flat, long bodies)
                    const float l1 = (rr + th) * (1.0f + 1e-6f) + 1e-6f;
                    if (box_dist2(cl, c_sc.obst_bmin[o], c_sc.obst_bmax[o]) > l1 * l1) continue;
                }
#pragma unroll 1
                for (int s = 0; s < cnt; ++s) {
                    const DevShape& ps = sm.shapes[off + s];
                    const V3 e = mk(cl.x - ps.cx, cl.y - ps.cy, cl.z - ps.cz);
                    const float l2 = rr + ps.radius + ps.margin + th;
                    const float l3 = (rr + ps.margin + th) * (1.0f + 1e-6f) + 1e-6f;   // against the box of the part's core
                    if (dot(e, e) <= l2 * l2 && box_dist2(cl, ps.bmin, ps.bmax) <= l3 * l3 &&
                        axis_lower_bound(sh, ps, F, TB) <= th) {
                        const int idx = atomicAdd(A.item_count, 1);
                        if (idx < A.capacity) write_item(A.items + idx, env, ia, off + s, GJK_CONTACT, sub, th, F, TB);
                        else atomicAdd(A.overflow, 1);
                        ++count;
                    }
                }
            }
        }
    }
    return count;
}

// Fine contact planning over the spans the coarse phase listed: one lane per sub-step of a span, 32 / span spans per
// warp (ten at 24 sub-steps).
#ifndef SM_CONTACT_MIN_BLOCKS
#define SM_CONTACT_MIN_BLOCKS 1   /* resident CTAs per SM the register allocation aims for (experiments: -DSM_CONTACT_MIN_BLOCKS=3) */
#endif
template <bool COUNT>
__global__ void __launch_bounds__(SM_WARPS_PER_BLOCK * 32, SM_CONTACT_MIN_BLOCKS) contact_plan_kernel(PlanArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n_units = A.cwork[0];
    const int span = (c_sc.substeps + SM_COARSE_LANES - 1) / SM_COARSE_LANES;
    const int upw = 32 / span;                                      // spans per warp: one lane per sub-step of a span
    if (blockIdx.x * SM_WARPS_PER_BLOCK * upw >= n_units) return;  // nothing for this block: skip the staging
    SmemLayout L = block_prologue(smem_raw, false);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const SceneSmem& sm = L.bs->scene;
    const int S = c_sc.substeps, stride = c_sc.contact_stride;
    const int kind = c_sc.n_obstacles > 0 ? c_sc.obst_kind[0] : SM_OBST_NONE;
    const double dt = xdiv(c_sc.ts, (double)S);
    unsigned n_flag = 0;
    const int lu = lane / span, i = lane - lu * span;
#pragma unroll 1
    for (int ub = (blockIdx.x * SM_WARPS_PER_BLOCK + warp) * upw; ub < n_units; ub += gridDim.x * SM_WARPS_PER_BLOCK * upw) {
        const int u = ub + lu;
        if (u >= n_units || lu >= upw) continue;
        const int unit = A.cwork[1 + u];
        const int env = unit / SM_COARSE_LANES, k = (unit % SM_COARSE_LANES) * span + i;  // 0-based sub-step
        if (k >= S || ((k + 1) % stride) != 0) continue;
        const float* scr = A.scratch + (size_t)env * SM_SCRATCH_FLOATS;
        const double* ob = A.obst + (size_t)env * SM_OBST_STRIDE;
        if (ob[SM_OB_LATCH] != 0.0) continue;
        const int idx0 = (int)ob[SM_OB_INDEX];
        bool test = false;
        int use_mask = 0;
        Xf T0, T1;
        xf_identity(T0);
        xf_identity(T1);
        if (kind == SM_OBST_PLANET && c_sc.terminate_moving) {
            planet_pose(0, (idx0 + k) % c_sc.planet_steps, T0);
            use_mask = 1;
            if (c_sc.n_obstacles > 1) { planet_pose(1, (idx0 + k) % c_sc.planet_steps, T1); use_mask = 3; }
            test = true;
        } else if (kind == SM_OBST_BALL && ob[SM_OB_BALL_ACTIVE] != 0.0) {
            const int sub = k + 1, k_end = ball_k_end(ob);
            if (sub <= k_end - 1) {
                // the test of sub-step `sub` happens after the ball moved to counter idx0+sub: the active area
                // uses the new position (ctlp.py:2851-2854), the manifold the previous one
                const double ball_t = ob[SM_OB_BALL_T];
                const double tn = ball_t + (double)sub * dt;
                const double px = ob[SM_OB_BALL_P0] + ob[SM_OB_BALL_V0] * tn;
                const double py = ob[SM_OB_BALL_P0 + 1] + ob[SM_OB_BALL_V0 + 1] * tn;
                if (sqrt(px * px + py * py) < c_sc.ball_active_xy) {
                    ball_pose(ob, ball_t + (double)k * dt, T0);
                    use_mask = 1;
                    test = true;
                }
            }
        }
        if (test) {
            const int c = contact_scan(sm, scr + k * SM_MAX_JOINTS, T0, T1, use_mask, A, env, k + 1);
            if (COUNT) n_flag += (unsigned)c;
        }
    }
    if (COUNT && A.counters && n_flag) atomicAdd(&A.counters[6], (unsigned long long)n_flag);
}

template <bool COUNT>
__global__ void __launch_bounds__(SM_WARPS_PER_BLOCK * 32) distance_plan_kernel(PlanArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemLayout L = block_prologue(smem_raw, false);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const SceneSmem& sm = L.bs->scene;
    WarpScratch& W = L.scratch[warp];
    const int S = c_sc.substeps;
    const int kind = c_sc.n_obstacles > 0 ? c_sc.obst_kind[0] : SM_OBST_NONE;
    const double dt = xdiv(c_sc.ts, (double)S);
    const float cap = (float)c_sc.static_cap, query = (float)c_sc.moving_query;
    const int n_fixed = c_sc.n_pairs_fixed, n_pairs = c_sc.n_pairs;
    const uint32_t* pair_tab = L.bs->pair_tab;
    const float2* pair_rm = L.bs->pair_rm;
    unsigned n_emit = 0;
#pragma unroll 1
    for (int env = blockIdx.x * SM_WARPS_PER_BLOCK + warp; env < A.n; env += gridDim.x * SM_WARPS_PER_BLOCK) {
        if (lane < SM_OBST_STRIDE) W.ob[lane] = A.obst[(size_t)env * SM_OBST_STRIDE + lane];
        const double q1 = A.kin[(size_t)env * SM_KIN_STRIDE + (lane & 7)];  // joint angle of the (new) knot
        __syncwarp();
        // ---------------- obstacle poses the reward distance is taken at: the end of the step if no contact latches
        // in it (a latched contact overrides the distance with 0 in finish_kernel, ctlp.py:3224-3234)
        bool moving = c_sc.n_mov_reward > 0 && W.ob[SM_OB_LATCH] == 0.0;
        if (kind == SM_OBST_PLANET) {
            const int idx = A.advance ? ((int)W.ob[SM_OB_INDEX] + S) % c_sc.planet_steps : (int)W.ob[SM_OB_INDEX];
            if (lane < c_sc.n_obstacles) planet_pose(lane, idx, W.obx[lane]);
        } else if (kind == SM_OBST_BALL) {
            double ball_t = W.ob[SM_OB_BALL_T];
            bool active = W.ob[SM_OB_BALL_ACTIVE] != 0.0;
            if (A.advance && active) {
                const int k_end = ball_k_end(W.ob);
                int adv = S;
                if (k_end <= S) { adv = k_end; active = false; }  // retired inside the step without hitting the robot
                ball_t += (double)adv * dt;   // pose for the float32 geometry only: the exact repeated addition of
                                              // self._t happens in finish_kernel, where the state is written
            }
            if (!active) moving = false;  // not in the list of observed obstacles (ctlp.py:3239-3245)
            if (lane == 0) ball_pose(W.ob, ball_t, W.obx[0]);
        } else if (kind == SM_OBST_HUMAN) {
            // the human at its setpoint pose (set_position_in_obstacle_client_to_setpoints, ctlp.py:3246-3251)
            human_fk_scan(A.hkin ? (float)A.hkin[(size_t)env * SM_KIN_STRIDE + (lane & 7)] : 0.0f, W.obx, lane);
        } else {
            moving = false;
        }
        frames_from_q32f(sm, (float)q1, W.fr, lane);
        if (lane < 4)   // slots 4.. belong to the human's braking-trajectory check
            A.res[(size_t)env * SM_RES_STRIDE + lane] = lane == 3 ? SM_RES_NO_CONTACT
                                                                   : fkey(lane == GJK_MOVING ? query + 0.002f : cap);
        __syncwarp();
        // ---------------- target point of the reaching task: was it reached in one of the S sub-steps?
        // (ctlp.py:2787-2821; target link point of the sub-step's setpoint pose)
        if (c_sc.use_target_points && A.target && A.advance) {
            double* tp = A.target + (size_t)env * SM_TP_STRIDE;
            const int nj = c_sc.n_joints;
            const V3 p1 = xf_apply(W.fr[nj], c_sc.tp_local[0], c_sc.tp_local[1], c_sc.tp_local[2]);
            bool reached = false;
            if (tp[SM_TP_ACTIVE] != 0.0) {
                const V3 T = mk((float)tp[SM_TP_POS], (float)tp[SM_TP_POS + 1], (float)tp[SM_TP_POS + 2]);
                const float* qs = A.scratch + (size_t)env * SM_SCRATCH_FLOATS + SM_QSET_OFF;
                // every sub-step's link point lies within infl of the new knot's: sum_j max_k |q_j(k) - q_j(S)| rho_j
                float infl = 1e-5f;
                const float q1f = (float)q1;
#pragma unroll 1
                for (int j = 0; j < nj; ++j) {
                    const float q1j = __shfl_sync(FULL, q1f, j);   // lane j holds joint j of the new knot
                    const float dq = lane < S ? fabsf(qs[lane * SM_MAX_JOINTS + j] - q1j) : 0.0f;
                    infl = fmaf(__uint_as_float(__reduce_max_sync(FULL, __float_as_uint(dq))), c_sc.tp_rho[j], infl);
                }
                const V3 e1 = p1 - T;
                if (sqrtf(dot(e1, e1)) - infl < (float)c_sc.tp_radius) {   // exact: lane k = sub-step k + 1
                    bool hit = false;
                    if (lane < S) {
                        Xf F;
                        xf_identity(F);
#pragma unroll 1
                        for (int j = 0; j < nj; ++j) fk_chain_step(sm, F, j, qs[lane * SM_MAX_JOINTS + j]);
                        const V3 e = xf_apply(F, c_sc.tp_local[0], c_sc.tp_local[1], c_sc.tp_local[2]) - T;
                        hit = sqrtf(dot(e, e)) < (float)c_sc.tp_radius;
                    }
                    reached = __any_sync(FULL, hit);
                }
            }
            if (lane == 0) {
                tp[SM_TP_LINK_POS] = (double)p1.x; tp[SM_TP_LINK_POS + 1] = (double)p1.y; tp[SM_TP_LINK_POS + 2] = (double)p1.z;
                tp[SM_TP_REACHED] = reached ? 1.0 : 0.0;
                if (reached) { tp[SM_TP_ACTIVE] = 0.0; tp[SM_TP_REACHED_N] += 1.0; }
            }
            __syncwarp();
        }
        // ---------------- world positions of every shape's sphere centre and centroid, once per env
#pragma unroll 1
        for (int sI = lane; sI < c_sc.n_shapes; sI += 32) {
            const DevShape& sh = sm.shapes[sI];
            const Xf& T = *frame_ptr(sh, W.fr, W.obx);
            const V3 c = xf_apply(T, sh.cx, sh.cy, sh.cz), g = xf_apply(T, sh.gx, sh.gy, sh.gz);
            W.pc[sI] = make_float4(c.x, c.y, c.z, 0.f);
            W.pg[sI] = make_float4(g.x, g.y, g.z, 0.f);
        }
        __syncwarp();
        // ---------------- pass 1 (lanes = pairs): the best upper bound of each class (distance of the hull centroids)
        const int np = moving ? n_pairs : n_fixed;
        float ub_s = cap, ub_e = (float)c_sc.self_query, ub_m = query;
        unsigned best_m = 0xffffffffu;  // the lane's moving pair with the smallest centroid distance
#pragma unroll 1
        for (int p = lane; p < np; p += 32) {
            const uint32_t e = pair_tab[p];
            const int ia = (int)(e & 0xfffu), ib = (int)((e >> 12) & 0xfffu), cls = (int)((e >> 24) & 3u);
            const float4 ga = W.pg[ia], gb = W.pg[ib];
            const V3 g = mk(ga.x - gb.x, ga.y - gb.y, ga.z - gb.z);
            const float ub = sqrtf(dot(g, g)) * (1.0f + 1e-6f) + 1e-7f - pair_rm[p].x;
            if (cls == GJK_STATIC) ub_s = fminf(ub_s, ub);
            else if (cls == GJK_SELF) ub_e = fminf(ub_e, ub);
            else if (ub < ub_m) { ub_m = ub; best_m = (unsigned)ia | ((unsigned)ib << 16); }
        }
        ub_s = funkey(__reduce_min_sync(FULL, fkey(ub_s)));
        ub_e = funkey(__reduce_min_sync(FULL, fkey(ub_e)));
        {   // moving class: tighten the centroid bound with the two support points of the closest-centroid pair along
            // the line between the centroids (two real points of the hulls: their distance bounds the minimum from
            // above); the fewer pairs stay below the bound, the fewer items the GJK kernel has to run
            const unsigned km = __reduce_min_sync(FULL, fkey(ub_m));
            const unsigned has = __ballot_sync(FULL, best_m != 0xffffffffu && fkey(ub_m) == km);
            ub_m = funkey(km);
            if (has) {
                const unsigned pr = __shfl_sync(FULL, best_m, __ffs(has) - 1);
                const DevShape& SA = sm.shapes[pr & 0xffffu];
                const DevShape& SB = sm.shapes[pr >> 16];
                const Xf& TA = *frame_ptr(SA, W.fr, W.obx);
                const Xf& TB = *frame_ptr(SB, W.fr, W.obx);
                const V3 dir = xf_apply(TB, SB.gx, SB.gy, SB.gz) - xf_apply(TA, SA.gx, SA.gy, SA.gz);  // A towards B
                const int sa = warp_support(c_sc.verts + SA.off, SA.cnt, xf_rot_t(TA, dir), lane);
                const int sb = warp_support(c_sc.verts + SB.off, SB.cnt, xf_rot_t(TB, mk(-dir.x, -dir.y, -dir.z)), lane);
                const float4 pa = __ldg(c_sc.verts + SA.off + sa), pb = __ldg(c_sc.verts + SB.off + sb);
                const V3 e = xf_apply(TA, pa.x, pa.y, pa.z) - xf_apply(TB, pb.x, pb.y, pb.z);
                ub_m = fminf(ub_m, sqrtf(dot(e, e)) * (1.0f + 1e-6f) + 1e-7f - SA.margin - SB.margin);
            }
        }
        // ---------------- pass 2: one item per pair whose lower bounds are below its class' upper bound.  The sphere
        // bound runs over all pairs (lanes = pairs); its survivors (a few per 32 pairs) are queued and the support-width
        // bound and the item write run on full batches of 32 survivors instead of on the few lanes of every batch
        unsigned short* Q = reinterpret_cast<unsigned short*>(W.fr2);   // [64]; fr2 is not used by this kernel
        int qn = 0;
        auto drain = [&](int cnt) {   // the first cnt (<= 32) queued pairs
            bool emit = false;
            float thr = 0.0f;
            int ia = 0, ib = 0, cls = 0;
            if (lane < cnt) {
                const uint32_t e = pair_tab[Q[lane]];
                ia = (int)(e & 0xfffu); ib = (int)((e >> 12) & 0xfffu); cls = (int)((e >> 24) & 3u);
                thr = cls == GJK_STATIC ? ub_s : cls == GJK_SELF ? ub_e : ub_m;
                const DevShape& SA = sm.shapes[ia];
                const DevShape& SB = sm.shapes[ib];
                emit = axis_lower_bound_d(SA, SB, *frame_ptr(SA, W.fr, W.obx), *frame_ptr(SB, W.fr, W.obx),
                                          mk(W.pc[ib].x - W.pc[ia].x, W.pc[ib].y - W.pc[ia].y, W.pc[ib].z - W.pc[ia].z)) <= thr;
            }
            const unsigned em = __ballot_sync(FULL, emit);
            if (em) {
                const int total = __popc(em);
                int b0 = 0;
                if (lane == 0) b0 = atomicAdd(A.item_count, total);
                b0 = __shfl_sync(FULL, b0, 0);
                if (b0 + total <= A.capacity) {
                    if (emit)
                        write_item(A.items + b0 + __popc(em & ((1u << lane) - 1u)), env, ia, ib, cls, 0, thr,
                                   *frame_ptr(sm.shapes[ia], W.fr, W.obx), *frame_ptr(sm.shapes[ib], W.fr, W.obx));
                } else if (lane == 0) {
                    atomicAdd(A.overflow, total);
                }
                if (COUNT) n_emit += (unsigned)total;
            }
        };
#pragma unroll 1
        for (int base = 0; base < np; base += 32) {
            const int p = base + lane;
            bool pass = false;
            if (p < np) {
                const uint32_t e = pair_tab[p];
                const int ia = (int)(e & 0xfffu), ib = (int)((e >> 12) & 0xfffu), cls = (int)((e >> 24) & 3u);
                const float thr = cls == GJK_STATIC ? ub_s : cls == GJK_SELF ? ub_e : ub_m;
                const float4 ca = W.pc[ia];
                if (e >> 31) {  // static shape in the world frame: sphere against its axis-aligned box
                    const DevShape& SB = sm.shapes[ib];
                    const float dx = fmaxf(fmaxf(SB.bmin[0] - ca.x, ca.x - SB.bmax[0]), 0.f);
                    const float dy = fmaxf(fmaxf(SB.bmin[1] - ca.y, ca.y - SB.bmax[1]), 0.f);
                    const float dz = fmaxf(fmaxf(SB.bmin[2] - ca.z, ca.z - SB.bmax[2]), 0.f);
                    pass = sqrtf(dx * dx + dy * dy + dz * dz) - pair_rm[p].y <= thr;
                } else {
                    const float4 cb = W.pc[ib];
                    const V3 d = mk(cb.x - ca.x, cb.y - ca.y, cb.z - ca.z);
                    pass = sqrtf(dot(d, d)) - pair_rm[p].y <= thr;
                }
            }
            const unsigned pm = __ballot_sync(FULL, pass);
            if (pm) {
                if (pass) Q[qn + __popc(pm & ((1u << lane) - 1u))] = (unsigned short)p;
                qn += __popc(pm);
                __syncwarp();
                if (qn >= 32) {
                    drain(32);
                    const unsigned short keep = Q[32 + lane];
                    __syncwarp();
                    Q[lane] = keep;
                    qn -= 32;
                    __syncwarp();
                }
            }
        }
        if (qn) drain(qn);
        __syncwarp();
    }
    if (COUNT && A.counters && lane == 0 && n_emit) atomicAdd(&A.counters[3], (unsigned long long)n_emit);
}

// the hook smenv_distances: results of the planned queries as the three distances (ctlp.py:3217-3374)
__global__ void distances_out_kernel(const unsigned* res, const double* obst, float* d_static, float* d_self,
                                     float* d_moving, int n) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= n) return;
    const unsigned* r = res + (size_t)env * SM_RES_STRIDE;
    d_static[env] = funkey(r[GJK_STATIC]);
    d_self[env] = funkey(r[GJK_SELF]);
    float dm = funkey(r[GJK_MOVING]);
    if (obst[(size_t)env * SM_OBST_STRIDE + SM_OB_LATCH] != 0.0 || dm <= 0.0f) dm = 0.0f;  // ctlp.py:3224-3234, :3277
    d_moving[env] = dm;
}
