// smenv_step.cuh -- geometry half of the env step, one WARP per env (SafeMotionsBase.step, safe_motions_base.py:
// 1043-1227): sub-step contacts, distances at the new knot, reward, termination, observation, auto-reset.
// The joint-space half (safe range, action mapping, setpoints) ran before in joint_kernel (smenv_joint.cuh).
#pragma once
#include "smenv_joint.cuh"

struct StepArgs {
    SmBuffers buf;
    int n;
    int auto_reset;
    uint32_t k0, k1;        // Philox key (seed)
    const float* scratch;   // [n][SM_SCRATCH_FLOATS] written by joint_kernel
    const double* start_pool;
    int start_pool_n;
    const double* ball_pool;
    int ball_pool_n;
    unsigned long long* counters;  // device SmCounters, or NULL
};

// Broad phase of the sub-step contact test: lane k checks sub-step k+1.  Returns a 2-bit mask (bit o = obstacle o)
// telling whether some robot bounding sphere comes within the contact threshold of the obstacle's bounding sphere at
// the poses Bullet's collision detection of that sub-step sees (tracked robot pose before integration, obstacle pose
// of the previous update).  Each lane runs its own serial FK chain.
__device__ __noinline__ int substep_broad_phase(const SceneSmem& sm, const float* qrow, V3 oc0, V3 oc1, int use_mask) {
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

template <bool COUNT>
__device__ void step_env(const StepArgs& A, int env, const float4* __restrict__ verts, WarpScratch& W, BlockShared& bs,
                         int lane) {
    const SceneSmem& sm = bs.scene;
    const int nj = c_sc.n_joints, S = c_sc.substeps;
    GjkCounters cnt = {0u, 0u, 0u, nullptr};
    GjkCounters* pc = COUNT ? &cnt : nullptr;
    long long tph = COUNT ? clock64() : 0;
    unsigned n_flagged = 0, n_contact_tests = 0;
#define SM_PHASE(i)                                                                          \
    if (COUNT) {                                                                             \
        long long now_ = clock64();                                                          \
        if (lane == 0) atomicAdd(&bs.counters[8 + (i)], (unsigned long long)(now_ - tph));   \
        tph = now_;                                                                          \
    }

    // ---------------- load the env records; the kinematic record already holds the new knot (joint_kernel)
    const double* kin = A.buf.kin + (size_t)env * SM_KIN_STRIDE;
    const float* scr = A.scratch + (size_t)env * SM_SCRATCH_FLOATS;
    const double q1 = kin[lane & 7];                                  // joint angle of the new knot
    if (lane < SM_OBST_STRIDE) W.ob[lane] = A.buf.obst[(size_t)env * SM_OBST_STRIDE + lane];
    for (int i = lane; i < S * SM_MAX_JOINTS; i += 32) (&W.qsub[0][0])[i] = scr[i];
    const int4 ep = *reinterpret_cast<const int4*>(A.buf.episode + 4 * (size_t)env);
    const int ep_len = ep.x + 1;  // safe_motions_base.py:1044
    const float rcode = scr[SM_MAX_SUB * SM_MAX_JOINTS + SM_MISC_RCODE];
    const float jerk_rel = scr[SM_MAX_SUB * SM_MAX_JOINTS + SM_MISC_JERK];
    const float umax = scr[SM_MAX_SUB * SM_MAX_JOINTS + SM_MISC_UMAX];
    const double dt = xdiv(c_sc.ts, (double)S);
    __syncwarp();
    SM_PHASE(0)

    // ---------------- sub-step contacts with the moving obstacles (ctlp.py:2590-2862)
    double latch = W.ob[SM_OB_LATCH];
    const int idx0 = (int)W.ob[SM_OB_INDEX];
    const int kind = c_sc.n_obstacles > 0 ? c_sc.obst_kind[0] : SM_OBST_NONE;
    const int stride = c_sc.contact_stride;
    int idx_new = idx0;
    double ball_t = W.ob[SM_OB_BALL_T], ball_active = W.ob[SM_OB_BALL_ACTIVE];

    if (kind == SM_OBST_PLANET) {
        if (stride > 0 && latch == 0.0 && c_sc.terminate_moving) {
            int f = 0;
            if (lane < S && ((lane + 1) % stride == 0)) {
                Xf T;
                planet_pose(0, (idx0 + lane) % c_sc.planet_steps, T);
                V3 oc0 = xf_apply(T, c_sc.obst_center[0][0], c_sc.obst_center[0][1], c_sc.obst_center[0][2]);
                V3 oc1 = oc0;
                if (c_sc.n_obstacles > 1) {
                    planet_pose(1, (idx0 + lane) % c_sc.planet_steps, T);
                    oc1 = xf_apply(T, c_sc.obst_center[1][0], c_sc.obst_center[1][1], c_sc.obst_center[1][2]);
                }
                f = substep_broad_phase(sm, W.qsub[lane], oc0, oc1, c_sc.n_obstacles > 1 ? 3 : 1);
            }
            unsigned m0 = __ballot_sync(FULL, f & 1), m1 = __ballot_sync(FULL, f & 2);
            unsigned m = m0 | m1;
            bool hit = false;
            n_flagged += __popc(m);
            SM_PHASE(1)
#pragma unroll 1
            while (m && !hit) {
                int k = __ffs(m) - 1;
                m &= m - 1;
                frames_from_q32(sm, W.qsub[k], W.fr2, lane);
                if (lane < c_sc.n_obstacles) planet_pose(lane, (idx0 + k) % c_sc.planet_steps, W.obx2[lane]);
                __syncwarp();
                if ((m0 >> k) & 1u) hit = contact_exists(verts, sm, 0, W.fr2, W.obx2, lane, pc);
                if (!hit && ((m1 >> k) & 1u)) hit = contact_exists(verts, sm, 1, W.fr2, W.obx2, lane, pc);
                __syncwarp();
            }
            if (hit) latch = 1.0;  // ctlp.py:2631-2637
        }
        idx_new = (idx0 + S) % c_sc.planet_steps;  // 24 x Planet.update (ctlp.py:4503-4505)
        if (lane < c_sc.n_obstacles) planet_pose(lane, idx_new, W.obx[lane]);
    } else if (kind == SM_OBST_BALL) {
        const double nmax = W.ob[SM_OB_BALL_NMAX], nhit = W.ob[SM_OB_BALL_NHIT];
        if (ball_active != 0.0) {
            // first sub-step (1-based) at which the ball leaves by the counters (ctlp.py:2840-2848, :4196-4211)
            int k1 = (int)nmax - idx0 + 1, k2 = (int)nhit - idx0;
            int k_end = k1 < k2 ? k1 : k2;
            if (k_end < 1) k_end = 1;
            int last_k = k_end - 1 < S ? k_end - 1 : S;  // sub-steps that still test contacts
            int kc = 0;                                   // sub-step of the first contact, 0 = none
            if (stride > 0 && latch == 0.0) {
                int f = 0;
                int sub = lane + 1;
                if (sub <= last_k && (sub % stride == 0)) {
                    // the test of sub-step `sub` happens after the ball moved to counter idx0+sub: the active area
                    // uses the new position (ctlp.py:2851-2854), the manifold the previous one
                    double tn = ball_t + (double)sub * dt;
                    double px = W.ob[SM_OB_BALL_P0] + W.ob[SM_OB_BALL_V0] * tn;
                    double py = W.ob[SM_OB_BALL_P0 + 1] + W.ob[SM_OB_BALL_V0 + 1] * tn;
                    if (sqrt(px * px + py * py) < c_sc.ball_active_xy) {
                        Xf T;
                        ball_pose(W.ob, ball_t + (double)lane * dt, T);
                        V3 oc0 = xf_apply(T, c_sc.obst_center[0][0], c_sc.obst_center[0][1], c_sc.obst_center[0][2]);
                        f = substep_broad_phase(sm, W.qsub[lane], oc0, oc0, 1);
                    }
                }
                unsigned m = __ballot_sync(FULL, f & 1);
                n_flagged += __popc(m);
                SM_PHASE(1)
#pragma unroll 1
                while (m && kc == 0) {
                    int k = __ffs(m) - 1;
                    m &= m - 1;
                    frames_from_q32(sm, W.qsub[k], W.fr2, lane);
                    if (lane == 0) ball_pose(W.ob, ball_t + (double)k * dt, W.obx2[0]);
                    __syncwarp();
                    if (contact_exists(verts, sm, 0, W.fr2, W.obx2, lane, pc)) kc = k + 1;
                    __syncwarp();
                }
            }
            int adv = S;
            if (kc > 0) { adv = kc; ball_active = 0.0; latch = 1.0; }       // hit robot (ctlp.py:2858-2861)
            else if (k_end <= S) { adv = k_end; ball_active = 0.0; }        // missed robot / hit obstacle
            idx_new = idx0 + adv;
#pragma unroll 1
            for (int i = 0; i < adv; ++i) ball_t = xadd(ball_t, dt);        // self._t += update_time_step
        }
        if (lane == 0) ball_pose(W.ob, ball_t, W.obx[0]);  // final obstacle pose for the reward distance
    }

    SM_PHASE(2)
    // ---------------- distances at the new knot (rewards.py:95-162; ctlp.py:3217-3374)
    frames_from_q64(sm, q1, W.fr, lane);
    __syncwarp();
    SM_PHASE(3)
    float d_static, d_self, d_moving;
    all_distances<COUNT>(verts, sm, W.fr, W.obx, latch != 0.0, kind == SM_OBST_BALL && ball_active == 0.0, d_static,
                         d_self, d_moving, lane, pc, bs.counters, tph);

    // ---------------- reward, termination (rewards.py:432-502; safe_motions_base.py:1775-1799): uniform float64
    double ds = (double)d_static, dse = (double)d_self, dm = (double)d_moving;
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
        atomicAdd(&bs.stats[0], 1.0);
        atomicAdd(&bs.stats[1], ep_return);
        atomicAdd(&bs.stats[2], (double)ep_len);
        atomicAdd(&bs.stats[3 + reason], 1.0);
    }

    SM_PHASE(6)
    // ---------------- new obstacle record; a ball that reached a final state is replaced when the observation is
    // taken (ctlp.py:2354-2360, :2893-2895).  The launch comes from the device-resident ball pool.
    int ball_draws = ep.z;
    __syncwarp();
    double ob_new = lane < SM_OBST_STRIDE ? W.ob[lane] : 0.0;
    if (lane == SM_OB_INDEX) ob_new = (double)idx_new;
    if (lane == SM_OB_LATCH) ob_new = latch;
    if (lane == SM_OB_BALL_T) ob_new = ball_t;
    if (lane == SM_OB_BALL_ACTIVE) ob_new = ball_active;
    if (kind == SM_OBST_BALL && ball_active == 0.0 && A.ball_pool_n > 0) {
        uint4 r = philox((uint32_t)env, (uint32_t)ball_draws, 0xBA11u, 0u, A.k0, A.k1);
        const double* e = A.ball_pool + (size_t)(r.x % (uint32_t)A.ball_pool_n) * SM_BALL_STRIDE;
        ball_draws++;
        if (lane >= SM_OB_BALL_P0 && lane < SM_OB_BALL_P0 + 10) ob_new = e[lane - SM_OB_BALL_P0];
        if (lane == SM_OB_INDEX || lane == SM_OB_BALL_T || lane == SM_OB_LATCH) ob_new = 0.0;
        if (lane == SM_OB_BALL_ACTIVE) ob_new = 1.0;
        if (lane == SM_OB_BALL_NMAX) ob_new = e[10];
        if (lane == SM_OB_BALL_NHIT) ob_new = e[11];
    }

    // ---------------- the kinematic record already holds the new knot; only an auto reset rewrites it
    int ep_len_new = ep_len, resets = ep.y;
    double ret_new = ep_return;
    const double* kin_obs = kin;
    if (done && A.auto_reset && A.start_pool_n > 0) {  // vector-env auto reset from the device-resident start pool
        uint4 r = philox((uint32_t)env, (uint32_t)resets, 0x5E7u, 1u, A.k0, A.k1);
        const double* e = A.start_pool + (size_t)(r.x % (uint32_t)A.start_pool_n) * SM_POOL_STRIDE;
        A.buf.kin[(size_t)env * SM_KIN_STRIDE + lane] = e[lane];
        if (lane < SM_OBST_STRIDE) ob_new = e[SM_KIN_STRIDE + lane];
        kin_obs = e;
        ep_len_new = 0;
        resets++;
        ret_new = 0.0;
    }
    if (lane < SM_OBST_STRIDE) {
        A.buf.obst[(size_t)env * SM_OBST_STRIDE + lane] = ob_new;
        W.ob[lane] = ob_new;
    }
    if (lane == 0) {
        *reinterpret_cast<int4*>(A.buf.episode + 4 * (size_t)env) = make_int4(ep_len_new, resets, ball_draws, ep.w);
        A.buf.ep_return[env] = ret_new;
    }
    __syncwarp();
    // ---------------- observation of the state the next action acts on (observations.py:313-351)
    write_observation(A.buf.obs + (size_t)env * c_sc.obs_size, kin_obs, W.ob, lane);
    __syncwarp();

    SM_PHASE(7)
    (void)n_contact_tests;
    if (COUNT && lane == 0) {
        atomicAdd(&bs.counters[6], (unsigned long long)n_flagged);
        atomicAdd(&bs.counters[0], (unsigned long long)cnt.calls);
        atomicAdd(&bs.counters[1], (unsigned long long)cnt.iters);
        atomicAdd(&bs.counters[2], (unsigned long long)cnt.dots);
        atomicAdd(&bs.counters[4], 1ull);
    }
}

template <bool COUNT>
__global__ void __launch_bounds__(SM_WARPS_PER_BLOCK * 32) step_kernel(StepArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemLayout L = block_prologue(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int env = blockIdx.x * SM_WARPS_PER_BLOCK + warp; env < A.n; env += gridDim.x * SM_WARPS_PER_BLOCK)
        step_env<COUNT>(A, env, L.verts, L.scratch[warp], *L.bs, lane);
    __syncthreads();
    if (A.buf.stats && tid < 16 && L.bs->stats[tid] != 0.0) atomicAdd(&A.buf.stats[tid], L.bs->stats[tid]);
    if (COUNT && A.counters && tid < 16 && L.bs->counters[tid] != 0ull) atomicAdd(&A.counters[tid], L.bs->counters[tid]);
}
