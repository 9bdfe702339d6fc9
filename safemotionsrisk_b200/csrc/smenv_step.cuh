// smenv_step.cuh -- geometry half of the env step as four PHASE KERNELS, one warp per env
// (SafeMotionsBase.step, safe_motions_base.py:1043-1227):
//
//   contact_broad_kernel   24 sub-steps on lanes: tracked-pose FK + bounding spheres -> which sub-steps need a
//                          narrow-phase contact test (ctlp.py:2590-2862)
//   contact_narrow_kernel  only the flagged envs (work list): GJK contact test of the flagged sub-steps
//   distance_kernel        obstacle bookkeeping, FK of the new knot, static / self / moving distances
//                          (rewards.py:95-162; ctlp.py:3217-3374)
//   finish_kernel          reward, termination, info, ball replacement, auto reset, observation
//                          (rewards.py:432-502; safe_motions_base.py:1775-1799; observations.py:313-351)
//
// The joint-space half (safe range, action mapping, setpoints) ran before in joint_kernel (smenv_joint.cuh).
// Why separate launches: the phases execute very different code; in one fused kernel the warps of an SM sat in
// different phases and the instruction cache thrashed (profiles/: stall_no_instruction 48 -> 7 -> ... cycles per issue).
// With one phase per launch every warp of an SM runs the same few KB of code.  The intermediates travel through a
// 1 KB per-env scratch record that stays in L2; HBM traffic is irrelevant for this FP32/latency-bound path.
#pragma once
#include "smenv_joint.cuh"

struct StepArgs {
    SmBuffers buf;
    int n;
    int auto_reset;
    uint32_t k0, k1;   // Philox key (seed)
    float* scratch;    // [n][SM_SCRATCH_FLOATS]
    int* worklist;     // [0] = count, [1..] = env indices flagged by the broad phase
    int* heavy;        // joint_heavy_kernel's list; its counter is cleared by finish_kernel for the next step
    const double* start_pool;
    int start_pool_n;
    const double* ball_pool;
    int ball_pool_n;
    unsigned long long* counters;  // device SmCounters, or NULL
};

// Broad phase of the sub-step contact test for ONE sub-step (one lane): own serial FK chain of the tracked pose, then
// robot bounding spheres against the obstacle bounding spheres, inflated by the contact thresholds.  Returns a 2-bit
// mask (bit o = obstacle o).  Poses are those Bullet's collision detection of that sub-step sees: tracked robot pose
// before integration, obstacle pose of the previous update.
__device__ __noinline__ int substep_broad_phase(const SceneSmem& sm, const float* __restrict__ qrow, V3 oc0, V3 oc1,
                                                int use_mask) {
    int hit = 0;
    Xf F;
    xf_identity(F);
#pragma unroll 1
    for (int f = 0; f <= c_sc.n_joints; ++f) {
        if (f > 0) {  // serial chain: frame f hangs off frame f-1 (checked on the host)
            const int j = f - 1;
            float s, c;
            sincosf(qrow[j], &s, &c);
            Xf L, C;
            float Rj[9];
            axis_angle(sm.jaxis[j][0], sm.jaxis[j][1], sm.jaxis[j][2], c, s, Rj);
            const float* A = sm.jR[j];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    L.r[3 * i + k] = fmaf(A[3 * i], Rj[k], fmaf(A[3 * i + 1], Rj[3 + k], A[3 * i + 2] * Rj[6 + k]));
            L.t[0] = sm.jt[j][0]; L.t[1] = sm.jt[j][1]; L.t[2] = sm.jt[j][2];
            xf_compose(F, L, C);
            F = C;
        }
#pragma unroll 1
        for (int slot = c_sc.contact_frame_start[f]; slot < c_sc.contact_frame_start[f + 1]; ++slot) {
            const DevShape& sh = sm.shapes[sm.mov_contact[slot]];
            V3 c = xf_apply(F, sh.cx, sh.cy, sh.cz);
            float rr = sh.radius + sh.margin;
            if (use_mask & 1) {
                V3 d = c - oc0;
                float lim = rr + c_sc.obst_radius[0] + sm.contact_thresh[0][slot];
                if (dot(d, d) <= lim * lim) hit |= 1;
            }
            if (use_mask & 2) {
                V3 d = c - oc1;
                float lim = rr + c_sc.obst_radius[1] + sm.contact_thresh[1][slot];
                if (dot(d, d) <= lim * lim) hit |= 2;
            }
        }
    }
    return hit;
}

// sub-steps (1-based) of a ball that still test contacts: all before the counters retire it (ctlp.py:2840-2848)
__device__ __forceinline__ int ball_k_end(const double* ob) {
    const int idx0 = (int)ob[SM_OB_INDEX];
    int k1 = (int)ob[SM_OB_BALL_NMAX] - idx0 + 1, k2 = (int)ob[SM_OB_BALL_NHIT] - idx0;
    int k_end = k1 < k2 ? k1 : k2;
    return k_end < 1 ? 1 : k_end;
}

// ------------------------------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(SM_WARPS_PER_BLOCK * 32) contact_broad_kernel(StepArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemLayout L = block_prologue(smem_raw, false);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const SceneSmem& sm = L.bs->scene;
    WarpScratch& W = L.scratch[warp];
    const int S = c_sc.substeps, stride = c_sc.contact_stride;
    const int kind = c_sc.n_obstacles > 0 ? c_sc.obst_kind[0] : SM_OBST_NONE;
    const double dt = xdiv(c_sc.ts, (double)S);
#pragma unroll 1
    for (int env = blockIdx.x * SM_WARPS_PER_BLOCK + warp; env < A.n; env += gridDim.x * SM_WARPS_PER_BLOCK) {
        float* scr = A.scratch + (size_t)env * SM_SCRATCH_FLOATS;
        if (lane < SM_OBST_STRIDE) W.ob[lane] = A.buf.obst[(size_t)env * SM_OBST_STRIDE + lane];
        __syncwarp();
        const int idx0 = (int)W.ob[SM_OB_INDEX];
        int f = 0;
        if (stride > 0 && W.ob[SM_OB_LATCH] == 0.0 && lane < S && ((lane + 1) % stride == 0)) {
            if (kind == SM_OBST_PLANET && c_sc.terminate_moving) {
                Xf T;
                planet_pose(0, (idx0 + lane) % c_sc.planet_steps, T);
                V3 oc0 = xf_apply(T, c_sc.obst_center[0][0], c_sc.obst_center[0][1], c_sc.obst_center[0][2]);
                V3 oc1 = oc0;
                if (c_sc.n_obstacles > 1) {
                    planet_pose(1, (idx0 + lane) % c_sc.planet_steps, T);
                    oc1 = xf_apply(T, c_sc.obst_center[1][0], c_sc.obst_center[1][1], c_sc.obst_center[1][2]);
                }
                f = substep_broad_phase(sm, scr + lane * SM_MAX_JOINTS, oc0, oc1, c_sc.n_obstacles > 1 ? 3 : 1);
            } else if (kind == SM_OBST_BALL && W.ob[SM_OB_BALL_ACTIVE] != 0.0) {
                const int sub = lane + 1, k_end = ball_k_end(W.ob);
                if (sub <= k_end - 1) {
                    // the test of sub-step `sub` happens after the ball moved to counter idx0+sub: the active area
                    // uses the new position (ctlp.py:2851-2854), the manifold the previous one
                    const double ball_t = W.ob[SM_OB_BALL_T];
                    double tn = ball_t + (double)sub * dt;
                    double px = W.ob[SM_OB_BALL_P0] + W.ob[SM_OB_BALL_V0] * tn;
                    double py = W.ob[SM_OB_BALL_P0 + 1] + W.ob[SM_OB_BALL_V0 + 1] * tn;
                    if (sqrt(px * px + py * py) < c_sc.ball_active_xy) {
                        Xf T;
                        ball_pose(W.ob, ball_t + (double)lane * dt, T);
                        V3 oc0 = xf_apply(T, c_sc.obst_center[0][0], c_sc.obst_center[0][1], c_sc.obst_center[0][2]);
                        f = substep_broad_phase(sm, scr + lane * SM_MAX_JOINTS, oc0, oc0, 1);
                    }
                }
            }
        }
        const unsigned m0 = __ballot_sync(FULL, f & 1), m1 = __ballot_sync(FULL, f & 2);
        if (lane == 0) {
            scr[SM_MISC_OFF + SM_MISC_MASK0] = __uint_as_float(m0);
            scr[SM_MISC_OFF + SM_MISC_MASK1] = __uint_as_float(m1);
            scr[SM_MISC_OFF + SM_MISC_HIT] = 0.0f;
            if (m0 | m1) A.worklist[1 + atomicAdd(A.worklist, 1)] = env;
            if (COUNT) atomicAdd(&L.bs->counters[6], (unsigned long long)__popc(m0 | m1));
        }
        __syncwarp();
    }
    __syncthreads();
    if (COUNT && A.counters && tid < 16 && L.bs->counters[tid] != 0ull) atomicAdd(&A.counters[tid], L.bs->counters[tid]);
}

// ------------------------------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(SM_WARPS_PER_BLOCK * 32) contact_narrow_kernel(StepArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n_work = A.worklist[0];
    if (blockIdx.x * SM_WARPS_PER_BLOCK >= n_work) return;  // nothing for this block: skip the staging
    SmemLayout L = block_prologue(smem_raw, true);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const SceneSmem& sm = L.bs->scene;
    WarpScratch& W = L.scratch[warp];
    const int S = c_sc.substeps;
    const int kind = c_sc.n_obstacles > 0 ? c_sc.obst_kind[0] : SM_OBST_NONE;
    const double dt = xdiv(c_sc.ts, (double)S);
    GjkCounters cnt = {0u, 0u, 0u, nullptr};
    GjkCounters* pc = COUNT ? &cnt : nullptr;
#pragma unroll 1
    for (int w = blockIdx.x * SM_WARPS_PER_BLOCK + warp; w < n_work; w += gridDim.x * SM_WARPS_PER_BLOCK) {
        const int env = A.worklist[1 + w];
        float* scr = A.scratch + (size_t)env * SM_SCRATCH_FLOATS;
        if (lane < SM_OBST_STRIDE) W.ob[lane] = A.buf.obst[(size_t)env * SM_OBST_STRIDE + lane];
        const unsigned m0 = __float_as_uint(scr[SM_MISC_OFF + SM_MISC_MASK0]);
        const unsigned m1 = __float_as_uint(scr[SM_MISC_OFF + SM_MISC_MASK1]);
        __syncwarp();
        const int idx0 = (int)W.ob[SM_OB_INDEX];
        const double ball_t = W.ob[SM_OB_BALL_T];
        unsigned m = m0 | m1;
        int kc = 0;  // 1-based sub-step of the first contact
#pragma unroll 1
        while (m && kc == 0) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            frames_from_q32(sm, scr + k * SM_MAX_JOINTS, W.fr2, lane);
            if (kind == SM_OBST_PLANET) {
                if (lane < c_sc.n_obstacles) planet_pose(lane, (idx0 + k) % c_sc.planet_steps, W.obx2[lane]);
            } else if (lane == 0) {
                ball_pose(W.ob, ball_t + (double)k * dt, W.obx2[0]);
            }
            __syncwarp();
            bool hit = false;
            if ((m0 >> k) & 1u) hit = contact_exists(L.verts, sm, 0, W.fr2, W.obx2, lane, pc);
            if (!hit && ((m1 >> k) & 1u)) hit = contact_exists(L.verts, sm, 1, W.fr2, W.obx2, lane, pc);
            if (hit) kc = k + 1;
            __syncwarp();
        }
        if (lane == 0) scr[SM_MISC_OFF + SM_MISC_HIT] = (float)kc;
        __syncwarp();
    }
    if (COUNT && lane == 0) {
        atomicAdd(&L.bs->counters[0], (unsigned long long)cnt.calls);
        atomicAdd(&L.bs->counters[1], (unsigned long long)cnt.iters);
        atomicAdd(&L.bs->counters[2], (unsigned long long)cnt.dots);
    }
    __syncthreads();
    if (COUNT && A.counters && tid < 16 && L.bs->counters[tid] != 0ull) atomicAdd(&A.counters[tid], L.bs->counters[tid]);
}

// ------------------------------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(SM_WARPS_PER_BLOCK * 32) distance_kernel(StepArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemLayout L = block_prologue(smem_raw, true);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const SceneSmem& sm = L.bs->scene;
    WarpScratch& W = L.scratch[warp];
    const int S = c_sc.substeps;
    const int kind = c_sc.n_obstacles > 0 ? c_sc.obst_kind[0] : SM_OBST_NONE;
    const double dt = xdiv(c_sc.ts, (double)S);
    GjkCounters cnt = {0u, 0u, 0u, nullptr};
    GjkCounters* pc = COUNT ? &cnt : nullptr;
#pragma unroll 1
    for (int env = blockIdx.x * SM_WARPS_PER_BLOCK + warp; env < A.n; env += gridDim.x * SM_WARPS_PER_BLOCK) {
        float* scr = A.scratch + (size_t)env * SM_SCRATCH_FLOATS;
        double* obg = A.buf.obst + (size_t)env * SM_OBST_STRIDE;
        if (lane < SM_OBST_STRIDE) W.ob[lane] = obg[lane];
        const double q1 = A.buf.kin[(size_t)env * SM_KIN_STRIDE + (lane & 7)];  // joint angle of the new knot
        const int kc = (int)scr[SM_MISC_OFF + SM_MISC_HIT];
        __syncwarp();
        // ---------------- obstacle bookkeeping of the 24 sub-steps (ctlp.py:2590-2862)
        double latch = W.ob[SM_OB_LATCH];
        const int idx0 = (int)W.ob[SM_OB_INDEX];
        int idx_new = idx0;
        double ball_t = W.ob[SM_OB_BALL_T], ball_active = W.ob[SM_OB_BALL_ACTIVE];
        if (kind == SM_OBST_PLANET) {
            if (kc > 0) latch = 1.0;                   // ctlp.py:2631-2637
            idx_new = (idx0 + S) % c_sc.planet_steps;  // 24 x Planet.update (ctlp.py:4503-4505)
            if (lane < c_sc.n_obstacles) planet_pose(lane, idx_new, W.obx[lane]);
        } else if (kind == SM_OBST_BALL) {
            if (ball_active != 0.0) {
                const int k_end = ball_k_end(W.ob);
                int adv = S;
                if (kc > 0) { adv = kc; ball_active = 0.0; latch = 1.0; }  // hit robot (ctlp.py:2858-2861)
                else if (k_end <= S) { adv = k_end; ball_active = 0.0; }   // missed robot / hit obstacle
                idx_new = idx0 + adv;
#pragma unroll 1
                for (int i = 0; i < adv; ++i) ball_t = xadd(ball_t, dt);   // self._t += update_time_step
            }
            if (lane == 0) ball_pose(W.ob, ball_t, W.obx[0]);  // final obstacle pose for the reward distance
        }
        if (lane == 0) {
            obg[SM_OB_INDEX] = (double)idx_new;
            obg[SM_OB_LATCH] = latch;
            if (kind == SM_OBST_BALL) { obg[SM_OB_BALL_T] = ball_t; obg[SM_OB_BALL_ACTIVE] = ball_active; }
        }
        // ---------------- distances at the new knot (rewards.py:95-162; ctlp.py:3217-3374)
        frames_from_q64(sm, q1, W.fr, lane);
        __syncwarp();
        float d_static, d_self, d_moving;
        all_distances(L.verts, sm, W.fr, W.obx, latch != 0.0, kind == SM_OBST_BALL && ball_active == 0.0, d_static,
                      d_self, d_moving, lane, pc);
        if (lane == 0) {
            scr[SM_MISC_OFF + SM_MISC_DSTATIC] = d_static;
            scr[SM_MISC_OFF + SM_MISC_DSELF] = d_self;
            scr[SM_MISC_OFF + SM_MISC_DMOVING] = d_moving;
        }
        __syncwarp();
    }
    if (COUNT && lane == 0) {
        atomicAdd(&L.bs->counters[0], (unsigned long long)cnt.calls);
        atomicAdd(&L.bs->counters[1], (unsigned long long)cnt.iters);
        atomicAdd(&L.bs->counters[2], (unsigned long long)cnt.dots);
    }
    __syncthreads();
    if (COUNT && A.counters && tid < 16 && L.bs->counters[tid] != 0ull) atomicAdd(&A.counters[tid], L.bs->counters[tid]);
}

// ------------------------------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(256) finish_kernel(StepArgs A) {
    __shared__ double s_stats[16];
    __shared__ double s_ob[8][SM_OBST_STRIDE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 16) s_stats[tid] = 0.0;
    if (blockIdx.x == 0 && tid == 0 && A.heavy) A.heavy[0] = 0;
    __syncthreads();
    const int env = blockIdx.x * 8 + warp;
    const int nj = c_sc.n_joints;
    const int kind = c_sc.n_obstacles > 0 ? c_sc.obst_kind[0] : SM_OBST_NONE;
    if (env < A.n) {
        const double* kin = A.buf.kin + (size_t)env * SM_KIN_STRIDE;
        const float* scr = A.scratch + (size_t)env * SM_SCRATCH_FLOATS + SM_MISC_OFF;
        double* ob = s_ob[warp];
        if (lane < SM_OBST_STRIDE) ob[lane] = A.buf.obst[(size_t)env * SM_OBST_STRIDE + lane];
        const int4 ep = *reinterpret_cast<const int4*>(A.buf.episode + 4 * (size_t)env);
        const int ep_len = ep.x + 1;  // safe_motions_base.py:1044
        const float rcode = (float)__float_as_int(scr[SM_MISC_RCODE]), jerk_rel = scr[SM_MISC_JERK], umax = scr[SM_MISC_UMAX];
        __syncwarp();
        const double latch = ob[SM_OB_LATCH], ball_active = ob[SM_OB_BALL_ACTIVE];

        // ---------------- reward, termination (rewards.py:432-502; safe_motions_base.py:1775-1799)
        double ds = (double)scr[SM_MISC_DSTATIC], dse = (double)scr[SM_MISC_DSELF], dm = (double)scr[SM_MISC_DMOVING];
        int c_static = 0, c_self = 0, c_moving = 0;
        if (ds < c_sc.collision_dist) { ds = 0.0; c_static = 1; }
        if (dse < c_sc.collision_dist) { dse = 0.0; c_self = 1; }
        if (dm < c_sc.collision_dist) { dm = 0.0; c_moving = 1; }
        double r_self = 0.0, r_static = 0.0, r_moving;
        if (c_sc.w_self != 0.0) { double r = fmin(1.0, dse / c_sc.d_self); r_self = r * r; }
        if (c_sc.w_static != 0.0) { double r = fmin(1.0, ds / c_sc.d_static); r_static = r * r; }
        { double r = fmin(1.0, dm / c_sc.d_moving); r_moving = r * r; }
        double action_punishment = 1.0;
        if (c_sc.punish_action) {
            double pu = ((double)umax - c_sc.action_thresh) / (1.0 - c_sc.action_thresh);  // rewards.py:18-21
            pu = fmax(0.0, fmin(1.0, pu));
            action_punishment = pu * pu;
        }
        double low_acc = 0.0, low_vel = 0.0;
        if (c_sc.w_low_acc != 0.0 || c_sc.w_low_vel != 0.0) {  // rewards.py:448-460
            const int j = lane & 7;
            float ra = lane < nj ? (float)fabs(kin[16 + j] / c_sc.acc_max[j]) : 0.0f;
            float rv = lane < nj ? (float)fabs(kin[8 + j] / c_sc.vel_max[j]) : 0.0f;
            ra = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(ra)));
            rv = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(rv)));
            double da = fmin(1.0, (double)ra / c_sc.thr_low_acc), dv = fmin(1.0, (double)rv / c_sc.thr_low_vel);
            if (c_sc.w_low_acc != 0.0) low_acc = (da - 1.0) * (da - 1.0);
            if (c_sc.w_low_vel != 0.0) low_vel = (dv - 1.0) * (dv - 1.0);
        }
        const bool term_coll = (c_sc.terminate_self && c_self) || (c_sc.terminate_static && c_static) ||
                               (c_sc.terminate_moving && c_moving);
        const bool finished = ep_len >= c_sc.episode_steps;  // trajectory_manager.py:187-192
        const double bonus = (finished && !term_coll) ? c_sc.termination_bonus : 0.0;
        const double punish = term_coll ? c_sc.early_termination_punishment : 0.0;
        const double reward = (1.0 - action_punishment) * c_sc.action_max_punishment + r_self * c_sc.w_self +
                              r_static * c_sc.w_static + r_moving * c_sc.w_moving + low_acc * c_sc.w_low_acc +
                              low_vel * c_sc.w_low_vel + bonus + punish;
        int done = 0, reason = SM_TERM_UNSET;
        if (c_sc.terminate_self && c_self) { done = 1; reason = SM_TERM_SELF_COLLISION; }
        else if (c_sc.terminate_static && c_static) { done = 1; reason = SM_TERM_STATIC_COLLISION; }
        else if (c_sc.terminate_moving && c_moving) { done = 1; reason = SM_TERM_MOVING_COLLISION; }
        else if (finished) { done = 1; reason = SM_TERM_TRAJECTORY_LENGTH; }
        const double ep_return = A.buf.ep_return[env] + reward;

        // ---------------- outputs of the finished step
        if (lane == 0) {
            A.buf.reward[env] = (float)reward;
            A.buf.done[env] = (uint8_t)done;
            A.buf.term_reason[env] = reason;
        }
        if (lane < SM_INFO_STRIDE) {
            float val = 0.0f;
            switch (lane) {
                case SM_INFO_D_STATIC: val = (float)ds; break;
                case SM_INFO_D_SELF: val = (float)dse; break;
                case SM_INFO_D_MOVING: val = (float)dm; break;
                case SM_INFO_COLL_STATIC: val = (float)c_static; break;
                case SM_INFO_COLL_SELF: val = (float)c_self; break;
                case SM_INFO_COLL_MOVING: val = (float)c_moving; break;
                case SM_INFO_ACTION_PUNISH: val = (float)action_punishment; break;
                case SM_INFO_R_STATIC: val = (float)r_static; break;
                case SM_INFO_R_SELF: val = (float)r_self; break;
                case SM_INFO_R_MOVING: val = (float)r_moving; break;
                case SM_INFO_EPISODE_LENGTH: val = (float)ep_len; break;
                case SM_INFO_EPISODE_RETURN: val = (float)ep_return; break;
                case SM_INFO_RANGE_CODE: val = rcode; break;
                case SM_INFO_CONTACT_LATCH: val = latch != 0.0 ? 1.0f : 0.0f; break;
                case SM_INFO_MAX_JERK_REL: val = jerk_rel; break;
                default: break;
            }
            A.buf.info[(size_t)env * SM_INFO_STRIDE + lane] = val;
        }
        if (lane == 0 && done) {  // episode statistics (train.py:59-117), aggregated per block first
            atomicAdd(&s_stats[0], 1.0);
            atomicAdd(&s_stats[1], ep_return);
            atomicAdd(&s_stats[2], (double)ep_len);
            atomicAdd(&s_stats[3 + reason], 1.0);
        }

        // ---------------- a ball that reached a final state is replaced when the observation is taken
        // (ctlp.py:2354-2360, :2893-2895); the launch comes from the device-resident ball pool
        int ball_draws = ep.z;
        bool ob_changed = false;
        double ob_new = lane < SM_OBST_STRIDE ? ob[lane] : 0.0;
        if (kind == SM_OBST_BALL && ball_active == 0.0 && A.ball_pool_n > 0) {
            uint4 r = philox((uint32_t)env, (uint32_t)ball_draws, 0xBA11u, 0u, A.k0, A.k1);
            const double* e = A.ball_pool + (size_t)(r.x % (uint32_t)A.ball_pool_n) * SM_BALL_STRIDE;
            ball_draws++;
            if (lane >= SM_OB_BALL_P0 && lane < SM_OB_BALL_P0 + 10) ob_new = e[lane - SM_OB_BALL_P0];
            if (lane == SM_OB_INDEX || lane == SM_OB_BALL_T || lane == SM_OB_LATCH) ob_new = 0.0;
            if (lane == SM_OB_BALL_ACTIVE) ob_new = 1.0;
            if (lane == SM_OB_BALL_NMAX) ob_new = e[10];
            if (lane == SM_OB_BALL_NHIT) ob_new = e[11];
            ob_changed = true;
        }
        // ---------------- vector-env auto reset from the device-resident start pool
        int ep_len_new = ep_len, resets = ep.y;
        double ret_new = ep_return;
        const double* kin_obs = kin;
        if (done && A.auto_reset && A.start_pool_n > 0) {
            uint4 r = philox((uint32_t)env, (uint32_t)resets, 0x5E7u, 1u, A.k0, A.k1);
            const double* e = A.start_pool + (size_t)(r.x % (uint32_t)A.start_pool_n) * SM_POOL_STRIDE;
            A.buf.kin[(size_t)env * SM_KIN_STRIDE + lane] = e[lane];
            if (lane < SM_OBST_STRIDE) ob_new = e[SM_KIN_STRIDE + lane];
            ob_changed = true;
            kin_obs = e;
            ep_len_new = 0;
            resets++;
            ret_new = 0.0;
        }
        __syncwarp();
        if (ob_changed && lane < SM_OBST_STRIDE) {
            A.buf.obst[(size_t)env * SM_OBST_STRIDE + lane] = ob_new;
            ob[lane] = ob_new;
        }
        if (lane == 0) {
            *reinterpret_cast<int4*>(A.buf.episode + 4 * (size_t)env) = make_int4(ep_len_new, resets, ball_draws, ep.w);
            A.buf.ep_return[env] = ret_new;
        }
        __syncwarp();
        // ---------------- observation of the state the next action acts on (observations.py:313-351)
        write_observation(A.buf.obs + (size_t)env * c_sc.obs_size, kin_obs, ob, lane);
    }
    __syncthreads();
    if (A.buf.stats && tid < 16 && s_stats[tid] != 0.0) atomicAdd(&A.buf.stats[tid], s_stats[tid]);
    if (COUNT && A.counters && tid == 0) {
        int first = blockIdx.x * 8;
        int cntb = A.n - first < 8 ? A.n - first : 8;
        if (cntb > 0) atomicAdd(&A.counters[4], (unsigned long long)cntb);
    }
}
