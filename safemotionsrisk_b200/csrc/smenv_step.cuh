// smenv_step.cuh -- one env step per warp (SafeMotionsBase.step, safe_motions_base.py:1043-1227).
#pragma once
#include "smenv_kernels.cuh"

struct StepArgs {
    SmBuffers buf;
    int n;
    int auto_reset;
    int random_actions;
    uint32_t k0, k1;        // Philox key (seed)
    uint32_t step_counter;  // launches so far: decorrelates the random actions of successive steps
    const double* start_pool;
    int start_pool_n;
    const double* ball_pool;
    int ball_pool_n;
    unsigned long long* counters;  // device SmCounters, or NULL
};

// Broad phase of the sub-step contact test: lane k checks sub-step k+1.  Returns, per obstacle, whether some robot
// bounding sphere comes within the contact threshold of the obstacle's bounding sphere at the poses that Bullet's
// collision detection of that sub-step sees (tracked robot pose before integration, obstacle pose of the previous
// update).
__device__ __forceinline__ void substep_broad_phase(const WarpScratch& W, int k, V3 oc0, V3 oc1, bool use0, bool use1,
                                                    bool& f0, bool& f1) {
    f0 = f1 = false;
    const float* qrow = W.qsub[k];
    Xf F;
#pragma unroll
    for (int i = 0; i < 9; ++i) F.r[i] = (i % 4 == 0) ? 1.0f : 0.0f;
    F.t[0] = F.t[1] = F.t[2] = 0.0f;
    for (int f = 0; f <= c_sc.n_joints; ++f) {
        if (f > 0) {  // serial chain: frame f hangs off frame f-1 (checked on the host)
            int j = f - 1;
            float s, c, R1[9], Rj[9];
            sincosf(qrow[j], &s, &c);
            mat_mul(F.r, c_sc.jR[j], R1);
            V3 tp = xf_apply(F, c_sc.jt[j][0], c_sc.jt[j][1], c_sc.jt[j][2]);
            axis_angle(c_sc.jaxis[j], c, s, Rj);
            mat_mul(R1, Rj, F.r);
            F.t[0] = tp.x; F.t[1] = tp.y; F.t[2] = tp.z;
        }
        for (int slot = c_sc.contact_frame_start[f]; slot < c_sc.contact_frame_start[f + 1]; ++slot) {
            const DevShape& sh = c_sc.shapes[c_sc.mov_contact[slot]];
            V3 c = xf_apply(F, sh.cx, sh.cy, sh.cz);
            float rr = sh.radius + sh.margin;
            if (use0) {
                V3 d = c - oc0;
                float lim = rr + c_sc.obst_radius[0] + c_sc.contact_thresh[0][slot];
                f0 = f0 || dot(d, d) <= lim * lim;
            }
            if (use1) {
                V3 d = c - oc1;
                float lim = rr + c_sc.obst_radius[1] + c_sc.contact_thresh[1][slot];
                f1 = f1 || dot(d, d) <= lim * lim;
            }
        }
    }
}

template <bool COUNT>
__device__ void step_env(const StepArgs& A, int env, const float4* __restrict__ verts, WarpScratch& W, BlockShared& bs,
                         int lane) {
    const int nj = c_sc.n_joints, S = c_sc.substeps;
    const int j = lane & 7;
    const bool jl = lane < nj;
    GjkCounters cnt = {0u, 0u, 0u};
    GjkCounters* pc = COUNT ? &cnt : nullptr;

    // ---------------- load the env records (coalesced: one 256-byte and one 128-byte row per env)
    double kv = A.buf.kin[(size_t)env * SM_KIN_STRIDE + lane];
    double q = shfl_d(kv, j), v = shfl_d(kv, 8 + j), a = shfl_d(kv, 16 + j), qa = shfl_d(kv, 24 + j);
    double ob = lane < SM_OBST_STRIDE ? A.buf.obst[(size_t)env * SM_OBST_STRIDE + lane] : 0.0;
    int4 ep = *reinterpret_cast<const int4*>(A.buf.episode + 4 * (size_t)env);
    const int ep_len = ep.x + 1;  // safe_motions_base.py:1044

    float uf = 0.0f;
    if (A.random_actions) {  // get_random_action (safe_motions_base.py:1327-1328), U(-1, 1) per joint
        uint4 r = philox((uint32_t)env, A.step_counter, (uint32_t)j, 0xAC71u, A.k0, A.k1);
        uf = 2.0f * u01f(r.x) - 1.0f;
    } else if (jl) {
        uf = A.buf.actions[(size_t)env * nj + j];
    }
    const double u = (double)uf;

    // ---------------- safe range, action mapping (lanes 0..nj-1, float64)
    double lo = 0.0, hi = 0.0, a1 = 0.0;
    int code = 0;
    if (jl) {
        safe_range_joint(j, q, v, a, lo, hi, code);
        a1 = map_action(u, lo, hi);
    }
    const int rcode = __reduce_or_sync(FULL, (unsigned)(jl ? code : 0));

    // ---------------- sub-steps: setpoints and the motor-tracked pose (safe_motions_base.py:1233-1277)
    const double dt = xdiv(c_sc.ts, (double)S);
    const double tvdt = xmul(c_sc.track_vel, dt);
    double q1 = q, v1 = v;
    if (jl) {
        for (int k = 1; k <= S; ++k) {
            double qs, vs, as;
            interpolate(q, v, a, a1, substep_time(k), qs, vs, as);
            W.qsub[k - 1][j] = (float)qa;  // pose seen by the collision detection of sub-step k
            qa = xadd(xadd(qa, xmul(c_sc.track_kp, xsub(qs, qa))), xmul(tvdt, vs));
            q1 = qs; v1 = vs;              // k == S leaves the new knot (safe_motions_base.py:1179-1181)
        }
    }
    float jerk_rel = jl ? (float)(fabs((a1 - a) / c_sc.ts) / c_sc.jerk_max[j]) : 0.0f;
    jerk_rel = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(jerk_rel)));  // non-negative floats
    __syncwarp();

    // ---------------- sub-step contacts with the moving obstacles (ctlp.py:2590-2862)
    double latch = shfl_d(ob, SM_OB_LATCH);
    const int idx0 = (int)shfl_d(ob, SM_OB_INDEX);
    const int kind = c_sc.n_obstacles > 0 ? c_sc.obst_kind[0] : SM_OBST_NONE;
    const int stride = c_sc.contact_stride;
    int idx_new = idx0;
    double ball_t = shfl_d(ob, SM_OB_BALL_T), ball_active = shfl_d(ob, SM_OB_BALL_ACTIVE);

    if (kind == SM_OBST_PLANET) {
        if (stride > 0 && latch == 0.0 && c_sc.terminate_moving) {
            bool f0 = false, f1 = false;
            if (lane < S && ((lane + 1) % stride == 0)) {
                Xf T;
                planet_pose(0, (idx0 + lane) % c_sc.planet_steps, T);
                V3 oc0 = xf_apply(T, c_sc.obst_center[0][0], c_sc.obst_center[0][1], c_sc.obst_center[0][2]);
                V3 oc1 = oc0;
                if (c_sc.n_obstacles > 1) {
                    planet_pose(1, (idx0 + lane) % c_sc.planet_steps, T);
                    oc1 = xf_apply(T, c_sc.obst_center[1][0], c_sc.obst_center[1][1], c_sc.obst_center[1][2]);
                }
                substep_broad_phase(W, lane, oc0, oc1, true, c_sc.n_obstacles > 1, f0, f1);
            }
            unsigned m0 = __ballot_sync(FULL, f0), m1 = __ballot_sync(FULL, f1);
            unsigned m = m0 | m1;
            bool hit = false;
            while (m && !hit) {
                int k = __ffs(m) - 1;
                m &= m - 1;
                frames_from_q32(W.qsub[k], W.fr2, lane);
                if (lane < c_sc.n_obstacles) planet_pose(lane, (idx0 + k) % c_sc.planet_steps, W.ob2[lane]);
                __syncwarp();
                if ((m0 >> k) & 1u) hit = contact_exists(verts, 0, W.fr2, W.ob2, lane, pc);
                if (!hit && ((m1 >> k) & 1u)) hit = contact_exists(verts, 1, W.fr2, W.ob2, lane, pc);
                __syncwarp();
            }
            if (hit) latch = 1.0;  // ctlp.py:2631-2637
        }
        idx_new = (idx0 + S) % c_sc.planet_steps;  // 24 x Planet.update (ctlp.py:4503-4505)
    } else if (kind == SM_OBST_BALL) {
        const double nmax = shfl_d(ob, SM_OB_BALL_NMAX), nhit = shfl_d(ob, SM_OB_BALL_NHIT);
        double bp0[3] = {shfl_d(ob, SM_OB_BALL_P0), shfl_d(ob, SM_OB_BALL_P0 + 1), shfl_d(ob, SM_OB_BALL_P0 + 2)};
        double bv0[3] = {shfl_d(ob, SM_OB_BALL_V0), shfl_d(ob, SM_OB_BALL_V0 + 1), shfl_d(ob, SM_OB_BALL_V0 + 2)};
        double be0[3] = {shfl_d(ob, SM_OB_BALL_EULER0), shfl_d(ob, SM_OB_BALL_EULER0 + 1),
                         shfl_d(ob, SM_OB_BALL_EULER0 + 2)};
        double bom = shfl_d(ob, SM_OB_BALL_OMEGA);
        if (ball_active != 0.0) {
            // first sub-step (1-based) at which the ball leaves by the counters (ctlp.py:2840-2848, :4196-4211)
            int k1 = (int)nmax - idx0 + 1, k2 = (int)nhit - idx0;
            int k_end = k1 < k2 ? k1 : k2;
            if (k_end < 1) k_end = 1;
            int last_k = k_end - 1 < S ? k_end - 1 : S;  // sub-steps that still test contacts
            int kc = 0;                                   // sub-step of the first contact, 0 = none
            if (stride > 0 && latch == 0.0) {
                bool f0 = false, f1 = false;
                int sub = lane + 1;
                if (sub <= last_k && (sub % stride == 0)) {
                    // the test of sub-step `sub` happens after the ball moved to counter idx0+sub: active area uses
                    // the new position (ctlp.py:2851-2854), the manifold the previous one
                    double tn = ball_t + (double)sub * dt;
                    double px = bp0[0] + bv0[0] * tn, py = bp0[1] + bv0[1] * tn;
                    if (sqrt(px * px + py * py) < c_sc.ball_active_xy) {
                        Xf T;
                        ball_pose(bp0, bv0, be0, bom, ball_t + (double)lane * dt, T);
                        V3 oc0 = xf_apply(T, c_sc.obst_center[0][0], c_sc.obst_center[0][1], c_sc.obst_center[0][2]);
                        substep_broad_phase(W, lane, oc0, oc0, true, false, f0, f1);
                    }
                }
                unsigned m = __ballot_sync(FULL, f0);
                while (m && kc == 0) {
                    int k = __ffs(m) - 1;
                    m &= m - 1;
                    frames_from_q32(W.qsub[k], W.fr2, lane);
                    if (lane == 0) ball_pose(bp0, bv0, be0, bom, ball_t + (double)k * dt, W.ob2[0]);
                    __syncwarp();
                    if (contact_exists(verts, 0, W.fr2, W.ob2, lane, pc)) kc = k + 1;
                    __syncwarp();
                }
            }
            int adv = S;
            if (kc > 0) { adv = kc; ball_active = 0.0; latch = 1.0; }       // hit robot (ctlp.py:2858-2861)
            else if (k_end <= S) { adv = k_end; ball_active = 0.0; }        // missed robot / hit obstacle
            idx_new = idx0 + adv;
            for (int i = 0; i < adv; ++i) ball_t = xadd(ball_t, dt);        // self._t += update_time_step
        }
        // ---------------- final obstacle pose for the reward distance
        if (lane == 0) ball_pose(bp0, bv0, be0, bom, ball_t, W.ob[0]);
    }
    if (kind == SM_OBST_PLANET && lane < c_sc.n_obstacles) planet_pose(lane, idx_new, W.ob[lane]);

    // ---------------- distances at the new knot (rewards.py:95-162; ctlp.py:3217-3374)
    frames_from_q64(q1, W.fr, lane);
    __syncwarp();
    const float cap = (float)c_sc.static_cap;
    float d_static = min_pair_list(verts, c_sc.static_pairs, c_sc.n_static_pairs, cap, W.fr, W.ob, lane, pc, nullptr);
    float d_self = min_pair_list(verts, c_sc.self_pairs, c_sc.n_self_pairs, cap, W.fr, W.ob, lane, pc, nullptr);
    const float query = (float)c_sc.moving_query;
    float d_moving = query + 0.002f;
    if (latch != 0.0) {
        d_moving = 0.0f;  // ctlp.py:3224-3234
    } else if (c_sc.n_mov_reward > 0) {
        for (int o = 0; o < c_sc.n_obstacles; ++o) {
            if (kind == SM_OBST_BALL && ball_active == 0.0) continue;
            d_moving = min_moving(verts, c_sc.mov_reward, c_sc.n_mov_reward, o, query, d_moving, W.fr, W.ob, lane, pc,
                                  nullptr);
            if (d_moving <= 0.0f) break;
        }
    }

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
        float mu = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(jl ? fabsf(uf) : 0.0f)));
        double pu = ((double)mu - c_sc.action_thresh) / (1.0 - c_sc.action_thresh);  // rewards.py:18-21
        pu = fmax(0.0, fmin(1.0, pu));
        action_punishment = pu * pu;
    }
    double low_acc = 0.0, low_vel = 0.0;
    if (c_sc.w_low_acc != 0.0 || c_sc.w_low_vel != 0.0) {  // rewards.py:448-460
        float ra = jl ? (float)fabs(a1 / c_sc.acc_max[j]) : 0.0f;
        float rv = jl ? (float)fabs(v1 / c_sc.vel_max[j]) : 0.0f;
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
            case SM_INFO_RANGE_CODE: val = (float)rcode; break;
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

    // ---------------- new obstacle record; a ball that reached a final state is replaced when the observation is
    // taken (ctlp.py:2354-2360, :2893-2895).  The launch comes from the device-resident ball pool.
    int ball_draws = ep.z;
    double ob_new = ob;
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

    // ---------------- new kinematic record (lane L: group L>>3 = q, v, a, q_act; joint L&7)
    double qq = shfl_d(q1, j), vv = shfl_d(v1, j), aa = shfl_d(a1, j), tt = shfl_d(qa, j);
    int grp = lane >> 3;
    double kin_new = j < nj ? (grp == 0 ? qq : grp == 1 ? vv : grp == 2 ? aa : tt) : 0.0;
    int ep_len_new = ep_len, resets = ep.y;
    double ret_new = ep_return;

    if (done && A.auto_reset && A.start_pool_n > 0) {  // vector-env auto reset from the device-resident start pool
        uint4 r = philox((uint32_t)env, (uint32_t)resets, 0x5E7u, 1u, A.k0, A.k1);
        const double* e = A.start_pool + (size_t)(r.x % (uint32_t)A.start_pool_n) * SM_POOL_STRIDE;
        kin_new = e[lane];
        if (lane < SM_OBST_STRIDE) ob_new = e[SM_KIN_STRIDE + lane];
        ep_len_new = 0;
        resets++;
        ret_new = 0.0;
    }
    A.buf.kin[(size_t)env * SM_KIN_STRIDE + lane] = kin_new;
    if (lane < SM_OBST_STRIDE) A.buf.obst[(size_t)env * SM_OBST_STRIDE + lane] = ob_new;
    if (lane == 0) {
        *reinterpret_cast<int4*>(A.buf.episode + 4 * (size_t)env) = make_int4(ep_len_new, resets, ball_draws, ep.w);
        A.buf.ep_return[env] = ret_new;
    }
    // ---------------- observation of the state the next action acts on (observations.py:313-351)
    write_observation(A.buf.obs + (size_t)env * c_sc.obs_size, shfl_d(kin_new, j), shfl_d(kin_new, 8 + j),
                      shfl_d(kin_new, 16 + j), ob_new, lane);

    if (COUNT && lane == 0) {
        atomicAdd(&bs.counters[0], (unsigned long long)cnt.calls);
        atomicAdd(&bs.counters[1], (unsigned long long)cnt.iters);
        atomicAdd(&bs.counters[2], (unsigned long long)cnt.dots);
        atomicAdd(&bs.counters[4], 1ull);
    }
}

template <bool COUNT>
__global__ void __launch_bounds__(SM_WARPS_PER_BLOCK * 32) step_kernel(StepArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* verts = reinterpret_cast<float4*>(smem_raw);
    size_t off = ((size_t)c_sc.n_verts * sizeof(float4) + 15) & ~(size_t)15;
    WarpScratch* scratch = reinterpret_cast<WarpScratch*>(smem_raw + off);
    BlockShared* bs = reinterpret_cast<BlockShared*>(smem_raw + off + SM_WARPS_PER_BLOCK * sizeof(WarpScratch));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < c_sc.n_verts; i += blockDim.x) verts[i] = __ldg(c_sc.verts + i);  // hulls -> shared memory
    if (tid < 16) bs->stats[tid] = 0.0;
    if (tid < 6) bs->counters[tid] = 0ull;
    __syncthreads();
    for (int env = blockIdx.x * SM_WARPS_PER_BLOCK + warp; env < A.n; env += gridDim.x * SM_WARPS_PER_BLOCK)
        step_env<COUNT>(A, env, verts, scratch[warp], *bs, lane);
    __syncthreads();
    if (A.buf.stats && tid < 16 && bs->stats[tid] != 0.0) atomicAdd(&A.buf.stats[tid], bs->stats[tid]);
    if (COUNT && A.counters && tid < 6 && bs->counters[tid] != 0ull) atomicAdd(&A.counters[tid], bs->counters[tid]);
}
