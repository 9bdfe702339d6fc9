// smenv_step.cuh -- last kernel of the env step, eight lanes per env (SafeMotionsBase.step, safe_motions_base.py:
// 1043-1227): obstacle bookkeeping of the 24 sub-steps, reward, termination, info, ball replacement, auto reset,
// observation (ctlp.py:2590-2862; rewards.py:432-502; safe_motions_base.py:1775-1799; observations.py:313-351).
//
// The step is a chain of small kernels, each with one job and a few KB of code (a fused version stalled ~48 cycles
// per issue on instruction fetch, profiles/):
//   joint_kernel / joint_heavy_kernel   safe range, action mapping, setpoints, tracked pose   (smenv_joint.cuh)
//   contact_plan_kernel                 sub-step broad phase -> contact items                  (smenv_plan.cuh)
//   distance_plan_kernel                FK of the new knot, sphere bounds -> distance items    (smenv_plan.cuh)
//   gjk_kernel                          one thread per item                                    (smenv_gjk.cuh)
//   finish_kernel                       this file
// Intermediates travel through per-env records that stay in L2 (scratch: sub-step poses; res: distance keys and the
// first contact sub-step).
#pragma once
#include "smenv_plan.cuh"

struct StepArgs {
    SmBuffers buf;
    int n;
    int auto_reset;
    uint32_t k0, k1;   // Philox key (seed)
    int env_base;      // index of env 0 of this launch in the whole vector env (the Philox counter uses the global index)
    float* scratch;    // [n][SM_SCRATCH_FLOATS]
    const unsigned* res;  // [n][SM_RES_STRIDE] written by distance_plan_kernel / gjk_kernel
    int* heavy;        // joint_heavy_kernel's list; its counter is cleared here for the next step
    const int* overflow;      // device flag of the planning kernels
    volatile int* host_flag;  // mapped pinned host word: set when the item buffer overflowed
    const double* start_pool;
    int start_pool_n;
    const double* ball_pool;
    int ball_pool_n;
    const double* target_pool;     // [target_pool_n][4] target points sampled on the device (reaching task)
    int target_pool_n;
    unsigned long long* counters;  // device SmCounters, or NULL
    const float* risk;             // [n] risk of the proposed action, or NULL when the gate is off
    const uint8_t* risky;          // [n] 1 where the gate replaced the action
};

// ------------------------------------------------------------------------------------------------------------------
// EIGHT LANES PER ENV, four envs per warp, 32 per block: almost everything here is scalar work per env (bookkeeping,
// reward algebra, termination), which a warp per env repeated on 32 lanes; the per-env records are 16 or 32 words, so
// a lane moves two or four of them.  Lane group g of a warp owns env 4 * warp_index + g; `sl` is the lane inside its
// group.  Groups past the last env run along on the last env with every store masked (`valid`), so that all lanes reach
// the warp-wide synchronisations.
#define SM_FINISH_ENVS_PER_BLOCK 32
template <bool COUNT>
#ifndef FIN_MIN_BLOCKS
#define FIN_MIN_BLOCKS 4   /* resident CTAs per SM the register allocation aims for: 64 registers instead of 78, 32 warps instead of
                              24 (space_bm step 680 -> 660 us, Human 1547 -> 1521) */
#endif
__global__ void __launch_bounds__(256, FIN_MIN_BLOCKS) finish_kernel(StepArgs A) {
    __shared__ double s_stats[16];
    __shared__ double s_ob[SM_FINISH_ENVS_PER_BLOCK][SM_OBST_STRIDE];
    const int tid = threadIdx.x, lane = tid & 31, sl = lane & 7;
    const unsigned gmask = 0xffu << (lane & 24);   // the lanes of this env's group
    if (tid < 16) s_stats[tid] = 0.0;
    if (blockIdx.x == 0 && tid == 0) {
        if (A.heavy) A.heavy[0] = 0;
        if (A.overflow && *A.overflow != 0 && A.host_flag) *A.host_flag = *A.overflow;
    }
    __syncthreads();
    const int env_raw = blockIdx.x * SM_FINISH_ENVS_PER_BLOCK + (tid >> 3);
    const bool valid = env_raw < A.n;
    const int env = valid ? env_raw : A.n - 1;
    const int nj = c_sc.n_joints;
    const int kind = c_sc.n_obstacles > 0 ? c_sc.obst_kind[0] : SM_OBST_NONE;
    {
        const double* kin = A.buf.kin + (size_t)env * SM_KIN_STRIDE;
        const float* scr = A.scratch + (size_t)env * SM_SCRATCH_FLOATS + SM_MISC_OFF;
        double* ob = s_ob[tid >> 3];
        static_assert(SM_OBST_STRIDE == 16 && SM_KIN_STRIDE == 32 && SM_INFO_STRIDE == 32, "record sizes of the 8-lane layout");
        ob[sl] = A.buf.obst[(size_t)env * SM_OBST_STRIDE + sl];
        ob[sl + 8] = A.buf.obst[(size_t)env * SM_OBST_STRIDE + sl + 8];
        const int4 ep = *reinterpret_cast<const int4*>(A.buf.episode + 4 * (size_t)env);
        const int ep_len = ep.x + 1;  // safe_motions_base.py:1044
        const float rcode = (float)__float_as_int(scr[SM_MISC_RCODE]), jerk_rel = scr[SM_MISC_JERK], umax = scr[SM_MISC_UMAX];
        const uint4 rs = *reinterpret_cast<const uint4*>(A.res + (size_t)env * SM_RES_STRIDE);
        const int kc = rs.w == SM_RES_NO_CONTACT ? 0 : (int)rs.w;  // 1-based sub-step of the first contact
        __syncwarp();
        // ---------------- obstacle bookkeeping of the S sub-steps (ctlp.py:2590-2862)
        double latch = ob[SM_OB_LATCH], ball_active = ob[SM_OB_BALL_ACTIVE];
        {
            const int S = c_sc.substeps;
            const int idx0 = (int)ob[SM_OB_INDEX];
            int idx_new = idx0;
            double ball_t = ob[SM_OB_BALL_T];
            if (kind == SM_OBST_HUMAN) {
                if (kc > 0) latch = 1.0;                   // Human._collision_detected (ctlp.py:4888-4898)
            } else if (kind == SM_OBST_PLANET) {
                if (kc > 0) latch = 1.0;                   // ctlp.py:2631-2637
                idx_new = (idx0 + S) % c_sc.planet_steps;  // S x Planet.update (ctlp.py:4503-4505)
            } else if (kind == SM_OBST_BALL && ball_active != 0.0) {
                const double dt = xdiv(c_sc.ts, (double)S);
                const int k_end = ball_k_end(ob);
                int adv = S;
                if (kc > 0) { adv = kc; ball_active = 0.0; latch = 1.0; }  // hit robot (ctlp.py:2858-2861)
                else if (k_end <= S) { adv = k_end; ball_active = 0.0; }   // missed robot / hit obstacle
                idx_new = idx0 + adv;
#pragma unroll 1
                for (int i = 0; i < adv; ++i) ball_t = xadd(ball_t, dt);   // self._t += update_time_step
            }
            __syncwarp();
            if (sl == 0) {
                ob[SM_OB_INDEX] = (double)idx_new;
                ob[SM_OB_LATCH] = latch;
                if (kind == SM_OBST_BALL) { ob[SM_OB_BALL_T] = ball_t; ob[SM_OB_BALL_ACTIVE] = ball_active; }
            }
            __syncwarp();
        }

        // ---------------- reward, termination (rewards.py:432-502; safe_motions_base.py:1775-1799)
        double ds = (double)funkey(rs.x), dse = (double)funkey(rs.y), dm = (double)funkey(rs.z);
        if (latch != 0.0 || dm <= 0.0) dm = 0.0;  // latched contact (ctlp.py:3224-3234), penetration (:3277-3278)
        int c_static = 0, c_self = 0, c_moving = 0;
        if (ds < c_sc.collision_dist) { ds = 0.0; c_static = 1; }
        if (dse < c_sc.collision_dist) { dse = 0.0; c_self = 1; }
        if (dm < c_sc.collision_dist) { dm = 0.0; c_moving = 1; }
        // distance rewards and action punishment in float32: their inputs are float32 (GJK distances, the action)
        // and the contract on rewards is 1e-4 relative, not bit-exactness (only the state is float64-exact)
        double r_self = 0.0, r_static = 0.0, r_moving;
        if (c_sc.w_self != 0.0) { float r = fminf(1.0f, (float)dse / (float)c_sc.d_self); r_self = (double)(r * r); }
        if (c_sc.w_static != 0.0) { float r = fminf(1.0f, (float)ds / (float)c_sc.d_static); r_static = (double)(r * r); }
        { float r = fminf(1.0f, (float)dm / (float)c_sc.d_moving); r_moving = (double)(r * r); }
        double action_punishment = 1.0;
        if (c_sc.punish_action) {
            float pu = (umax - (float)c_sc.action_thresh) / (1.0f - (float)c_sc.action_thresh);  // rewards.py:18-21
            pu = fmaxf(0.0f, fminf(1.0f, pu));
            action_punishment = (double)(pu * pu);
        }
        double low_acc = 0.0, low_vel = 0.0;
        if (c_sc.w_low_acc != 0.0 || c_sc.w_low_vel != 0.0) {  // rewards.py:448-460
            float ra = sl < nj ? (float)fabs(kin[16 + sl] / c_sc.acc_max[sl]) : 0.0f;
            float rv = sl < nj ? (float)fabs(kin[8 + sl] / c_sc.vel_max[sl]) : 0.0f;
            ra = __uint_as_float(__reduce_max_sync(gmask, __float_as_uint(ra)));
            rv = __uint_as_float(__reduce_max_sync(gmask, __float_as_uint(rv)));
            double da = fmin(1.0, (double)ra / c_sc.thr_low_acc), dv = fmin(1.0, (double)rv / c_sc.thr_low_vel);
            if (c_sc.w_low_acc != 0.0) low_acc = (da - 1.0) * (da - 1.0);
            if (c_sc.w_low_vel != 0.0) low_vel = (dv - 1.0) * (dv - 1.0);
        }
        const bool term_coll = (c_sc.terminate_self && c_self) || (c_sc.terminate_static && c_static) ||
                               (c_sc.terminate_moving && c_moving);
        const bool finished = ep_len >= c_sc.episode_steps;  // trajectory_manager.py:187-192
        const double bonus = (finished && !term_coll) ? c_sc.termination_bonus : 0.0;
        const double punish = term_coll ? c_sc.early_termination_punishment : 0.0;
        double reward = (1.0 - action_punishment) * c_sc.action_max_punishment + r_self * c_sc.w_self +
                        r_static * c_sc.w_static + r_moving * c_sc.w_moving + low_acc * c_sc.w_low_acc +
                        low_vel * c_sc.w_low_vel + bonus + punish;
        // ---------------- reaching task: TargetPointReachingReward (rewards.py:303-396; ctlp.py:2309-2350)
        double* tp = (c_sc.use_target_points && A.buf.target) ? A.buf.target + (size_t)env * SM_TP_STRIDE : nullptr;
        double tp_reward = 0.0;
        bool tp_reached = false;
        if (tp) {
            const bool active = tp[SM_TP_ACTIVE] != 0.0;
            tp_reached = tp[SM_TP_REACHED] != 0.0;
            double norm = 1.0;
            if ((active || tp_reached) && c_sc.tp_normalize) {
                norm = tp[SM_TP_INIT_DIST];
                if (norm == 0.0) norm += 0.0000001;
            }
            if (active) {
                const double dx = tp[SM_TP_POS] - tp[SM_TP_LINK_POS], dy = tp[SM_TP_POS + 1] - tp[SM_TP_LINK_POS + 1],
                             dz = tp[SM_TP_POS + 2] - tp[SM_TP_LINK_POS + 2];
                tp_reward = (tp[SM_TP_LAST_DIST] - sqrt(dx * dx + dy * dy + dz * dz)) / (c_sc.ts * norm);
            } else if (tp_reached) {
                tp_reward = tp[SM_TP_LAST_DIST] / (c_sc.ts * norm) + c_sc.tp_bonus;
            }
            const double pun = c_sc.punish_action ? action_punishment : 0.0;
            reward = tp_reward * c_sc.tp_reward_factor - pun * c_sc.action_max_punishment + r_self * c_sc.w_self +
                     r_static * c_sc.w_static + r_moving * c_sc.w_moving;
        }
        int done = 0, reason = SM_TERM_UNSET;
        if (c_sc.terminate_self && c_self) { done = 1; reason = SM_TERM_SELF_COLLISION; }
        else if (c_sc.terminate_static && c_static) { done = 1; reason = SM_TERM_STATIC_COLLISION; }
        else if (c_sc.terminate_moving && c_moving) { done = 1; reason = SM_TERM_MOVING_COLLISION; }
        else if (finished) { done = 1; reason = SM_TERM_TRAJECTORY_LENGTH; }
        const double reward_raw = reward;
        reward *= c_sc.reward_scale;   // normalize_reward_to_frequency (rewards.py:172-176)
        const double ep_return = A.buf.ep_return[env] + reward;
        // risk gate bookkeeping (actions.py:335-340): ep.w = 1 + first risky step of the episode (0 = none so far)
        const int risky_now = A.risky ? (int)A.risky[env] : 0;
        int first_risky = ep.w;
        if (risky_now && first_risky == 0) first_risky = ep_len;   // episode_length - 1 of the reference, plus one
        __syncwarp();   // every lane has read the target-point record and the running return before lane 0 updates them

        // ---------------- outputs of the finished step
        if (sl == 0 && valid) {
            A.buf.reward[env] = (float)reward;
            A.buf.done[env] = (uint8_t)done;
            A.buf.term_reason[env] = reason;
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int slot = sl + 8 * h;
            float val = 0.0f;
            switch (slot) {
                case SM_INFO_RISKY_ACTION: val = (float)risky_now; break;
                case SM_INFO_RISK: val = A.risk ? A.risk[env] : 0.0f; break;
                case SM_INFO_FIRST_RISKY_STEP: val = (float)(first_risky - 1); break;
                case SM_INFO_REWARD_RAW: val = (float)reward_raw; break;
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
                case SM_INFO_TP_REWARD: val = (float)tp_reward; break;
                default: break;
            }
            if (valid) A.buf.info[(size_t)env * SM_INFO_STRIDE + slot] = val;
        }
        // ---------------- per-episode aggregation of the step info (train.py:59-117: average / max / min over the episode)
        if (A.buf.epacc) {
            // kinematic part of the observation before clipping (observations.py:365-413)
            float prel = 0.f, vrel = 0.f, arel = 0.f;
            if (sl < nj) {
                prel = (float)normalize_mm(kin[sl], c_sc.pos_lo[sl], c_sc.pos_hi[sl]);
                vrel = (float)(kin[8 + sl] / c_sc.vel_max[sl]);
                arel = (float)(kin[16 + sl] / c_sc.acc_max[sl]);
            }
            float vn2 = vrel * vrel;
            float pm = fabsf(prel), vm = fabsf(vrel), am = fabsf(arel);
#pragma unroll
            for (int m = 1; m < 8; m <<= 1) {
                vn2 += __shfl_xor_sync(FULL, vn2, m);
                pm = fmaxf(pm, __shfl_xor_sync(FULL, pm, m));
                vm = fmaxf(vm, __shfl_xor_sync(FULL, vm, m));
                am = fmaxf(am, __shfl_xor_sync(FULL, am, m));
            }
            float* acc = A.buf.epacc + (size_t)env * SM_EP_STRIDE;
            const bool first_step = ep_len == 1;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = sl + 8 * h;
                float x = 0.f;
                switch (k) {
                    case SM_EP_COLL_SELF: x = (float)c_self; break;
                    case SM_EP_COLL_STATIC: x = (float)c_static; break;
                    case SM_EP_COLL_MOVING: x = (float)c_moving; break;
                    case SM_EP_ACTION_PUNISH: x = c_sc.punish_action || !tp ? (float)action_punishment : 0.f; break;
                    case SM_EP_R_SELF: x = (float)r_self; break;
                    case SM_EP_R_STATIC: x = (float)r_static; break;
                    case SM_EP_R_MOVING: x = (float)r_moving; break;
                    case SM_EP_REWARD: x = (float)reward; break;
                    case SM_EP_TP_REWARD: x = (float)tp_reward; break;
                    case SM_EP_RISKY_ACTION: x = (float)risky_now; break;
                    case SM_EP_JOINT_VEL_NORM: x = sqrtf(vn2); break;
                    case SM_EP_POS_VIOLATION: x = pm > 1.001f ? 1.f : 0.f; break;
                    case SM_EP_VEL_VIOLATION: x = vm > 1.001f ? 1.f : 0.f; break;
                    case SM_EP_ACC_VIOLATION: x = am > 1.001f ? 1.f : 0.f; break;
                    case SM_EP_JERK_VIOLATION: x = jerk_rel > 1.002f ? 1.f : 0.f; break;
                    case SM_EP_OBS_CLIPPING: x = (pm > 1.f || vm > 1.f || am > 1.f) ? 1.f : 0.f; break;
                    default: break;
                }
                float s0 = first_step ? 0.f : acc[k], mx = first_step ? -FLT_MAX : acc[16 + k], mn = first_step ? FLT_MAX : acc[32 + k];
                s0 += x; mx = fmaxf(mx, x); mn = fminf(mn, x);
                float* dst = (done && valid) ? A.buf.epinfo + (size_t)env * SM_EP_STRIDE : acc;
                if (valid) { dst[k] = s0; dst[16 + k] = mx; dst[32 + k] = mn; }
            }
            // episode counters: one lane each
            if (valid && sl < 8) {
                const int slot = 48 + sl;
                float cprev = first_step ? 0.f : acc[slot], cnew = cprev;
                switch (slot) {
                    case SM_EPC_BALLS_HIT_ROBOT: cnew = cprev + ((kind == SM_OBST_BALL && kc > 0) ? 1.f : 0.f); break;
                    case SM_EPC_BALLS_MISSED:
                        cnew = cprev + ((kind == SM_OBST_BALL && kc == 0 && ob[SM_OB_BALL_ACTIVE] == 0.0 && ball_active == 0.0 &&
                                         A.buf.obst[(size_t)env * SM_OBST_STRIDE + SM_OB_BALL_ACTIVE] != 0.0) ? 1.f : 0.f);
                        break;
                    case SM_EPC_TARGETS_REACHED: cnew = cprev + (tp_reached ? 1.f : 0.f); break;
                    case SM_EPC_FIRST_RISKY_STEP: cnew = (float)(first_risky - 1); break;
                    case SM_EPC_LENGTH: cnew = (float)ep_len; break;
                    case SM_EPC_RETURN: cnew = (float)ep_return; break;
                    case SM_EPC_REASON: cnew = (float)reason; break;
                    case SM_EPC_HUMAN_BRAKED:
                        cnew = cprev + ((A.buf.hstate && A.buf.hstate[(size_t)env * SM_HSTATE_STRIDE + SM_HS_BRAKED] != 0.0) ? 1.f : 0.f);
                        break;
                    default: break;
                }
                float* dst = done ? A.buf.epinfo + (size_t)env * SM_EP_STRIDE : acc;
                dst[slot] = cnew;
            }
        }
        if (sl == 0 && done && valid) {  // episode statistics (train.py:59-117), aggregated per block first
            atomicAdd(&s_stats[0], 1.0);
            atomicAdd(&s_stats[1], ep_return);
            atomicAdd(&s_stats[2], (double)ep_len);
            atomicAdd(&s_stats[3 + reason], 1.0);
        }

        // ---------------- a ball that reached a final state is replaced when the observation is taken
        // (ctlp.py:2354-2360, :2893-2895); the launch comes from the device-resident ball pool
        int ball_draws = ep.z;
        double ob_new[2] = {ob[sl], ob[sl + 8]};   // the lane's two words of the obstacle record
        if (kind == SM_OBST_BALL && ball_active == 0.0 && A.ball_pool_n > 0) {
            uint4 r = philox((uint32_t)(env + A.env_base), (uint32_t)ball_draws, 0xBA11u, 0u, A.k0, A.k1);
            const double* e = A.ball_pool + (size_t)(r.x % (uint32_t)A.ball_pool_n) * SM_BALL_STRIDE;
            ball_draws++;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int w = sl + 8 * h;
                if (w >= SM_OB_BALL_P0 && w < SM_OB_BALL_P0 + 10) ob_new[h] = e[w - SM_OB_BALL_P0];
                if (w == SM_OB_INDEX || w == SM_OB_BALL_T || w == SM_OB_LATCH) ob_new[h] = 0.0;
                if (w == SM_OB_BALL_ACTIVE) ob_new[h] = 1.0;
                if (w == SM_OB_BALL_NMAX) ob_new[h] = e[10];
                if (w == SM_OB_BALL_NHIT) ob_new[h] = e[11];
            }
        }
        // ---------------- vector-env auto reset from the device-resident start pool
        int ep_len_new = ep_len, resets = ep.y;
        double ret_new = ep_return;
        const double* kin_obs = kin;
        if (done && A.auto_reset && A.start_pool_n > 0) {
            uint4 r = philox((uint32_t)(env + A.env_base), (uint32_t)resets, 0x5E7u, 1u, A.k0, A.k1);
            const double* e = A.start_pool + (size_t)(r.x % (uint32_t)A.start_pool_n) * SM_POOL_STRIDE;
            if (valid) {
#pragma unroll
                for (int h = 0; h < 4; ++h) A.buf.kin[(size_t)env * SM_KIN_STRIDE + sl + 8 * h] = e[sl + 8 * h];
            }
            ob_new[0] = e[SM_KIN_STRIDE + sl];
            ob_new[1] = e[SM_KIN_STRIDE + sl + 8];
            kin_obs = e;
            ep_len_new = 0;
            resets++;
            ret_new = 0.0;
            first_risky = 0;
        }
        __syncwarp();
        const bool was_reset = kin_obs != kin;
        if (tp) {  // get_target_point_observation (ctlp.py:2210-2271): replace a reached point, record the distances
            if (sl == 0 && valid) {
                if (was_reset) {
                    target_episode_start(tp, kin_obs, nullptr, A.target_pool, A.target_pool_n, env + A.env_base, A.k0, A.k1);
                } else {
                    if (tp_reached && A.target_pool_n > 0) {
                        uint4 r = philox((uint32_t)(env + A.env_base), (uint32_t)tp[SM_TP_DRAWS], 0x7A26u, 2u, A.k0, A.k1);
                        const double* e = A.target_pool + (size_t)(r.x % (uint32_t)A.target_pool_n) * 4;
                        tp[SM_TP_POS] = e[0]; tp[SM_TP_POS + 1] = e[1]; tp[SM_TP_POS + 2] = e[2];
                        tp[SM_TP_ACTIVE] = 1.0;
                        tp[SM_TP_INIT_DIST] = nan("");
                        tp[SM_TP_DRAWS] += 1.0;
                    }
                    tp[SM_TP_REACHED] = 0.0;
                    if (tp[SM_TP_ACTIVE] != 0.0) {
                        const double dx = tp[SM_TP_POS] - tp[SM_TP_LINK_POS], dy = tp[SM_TP_POS + 1] - tp[SM_TP_LINK_POS + 1],
                                     dz = tp[SM_TP_POS + 2] - tp[SM_TP_LINK_POS + 2];
                        tp[SM_TP_LAST_DIST] = sqrt(dx * dx + dy * dy + dz * dz);
                        if (isnan(tp[SM_TP_INIT_DIST])) tp[SM_TP_INIT_DIST] = tp[SM_TP_LAST_DIST];
                    }
                }
                __threadfence_block();
            }
            __syncwarp();
        }
        ob[sl] = ob_new[0];
        ob[sl + 8] = ob_new[1];
        if (valid) {
            A.buf.obst[(size_t)env * SM_OBST_STRIDE + sl] = ob_new[0];
            A.buf.obst[(size_t)env * SM_OBST_STRIDE + sl + 8] = ob_new[1];
            if (sl == 0) {
                *reinterpret_cast<int4*>(A.buf.episode + 4 * (size_t)env) = make_int4(ep_len_new, resets, ball_draws, first_risky);
                A.buf.ep_return[env] = ret_new;
            }
        }
        __syncwarp();
        // ---------------- observation of the state the next action acts on (observations.py:313-351)
        if (valid)
            write_observation(A.buf.obs + (size_t)env * c_sc.obs_size, kin_obs, ob, tp, sl, 8,
                              A.buf.hobs ? A.buf.hobs + (size_t)env * SM_HOBS_STRIDE : nullptr);
    }
    __syncthreads();
    if (A.buf.stats && tid < 16 && s_stats[tid] != 0.0) atomicAdd(&A.buf.stats[tid], s_stats[tid]);
    if (COUNT && A.counters && tid == 0) {
        int first = blockIdx.x * SM_FINISH_ENVS_PER_BLOCK;
        int cntb = A.n - first < SM_FINISH_ENVS_PER_BLOCK ? A.n - first : SM_FINISH_ENVS_PER_BLOCK;
        if (cntb > 0) atomicAdd(&A.counters[4], (unsigned long long)cntb);
    }
}
