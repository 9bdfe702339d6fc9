// smenv_kernels.cuh -- per-warp scratch, obstacle poses, pair-list minima, contact test, observation.
#pragma once
#include "smenv_geom.cuh"

#define SM_WARPS_PER_BLOCK 8
#define SM_POOL_STRIDE 48  /* doubles per start-pool entry: kin record (32) + obstacle record (16) */
#define SM_BALL_STRIDE 12  /* doubles per ball-pool entry: p0 v0 euler0 omega nmax nhit */
#define SM_MAX_SUB 32

// per-warp scratch in shared memory
struct WarpScratch {
    double ob[SM_OBST_STRIDE];              // the env's obstacle record (broadcast reads instead of shuffles)
    Xf fr[1 + SM_MAX_JOINTS];               // robot frames at the end-of-step setpoint pose
    Xf fr2[1 + SM_MAX_JOINTS];              // robot frames of one sub-step (narrow phase)
    Xf obx[SM_MAX_OBST_FRAMES];             // obstacle frames at the end of the step (human: base + 8 joint frames)
    Xf obx2[SM_MAX_OBST_FRAMES];            // obstacle frames of one sub-step
    float4 pc[SM_MAX_SHAPES];               // world position of every shape's bounding-sphere centre (distance planning)
    float4 pg[SM_MAX_SHAPES];               // world position of every shape's centroid
};

struct BlockShared {
    SceneSmem scene;                        // \ the first three members mirror SceneImage
    uint32_t pair_tab[SM_MAX_PLAN_PAIRS];   // |
    float2 pair_rm[SM_MAX_PLAN_PAIRS];      // /
    double stats[16];
    unsigned long long counters[16];
};

// dynamic shared memory layout of the geometry kernels: hull vertices | BlockShared | WarpScratch[warps]
struct SmemLayout {
    float4* verts;
    BlockShared* bs;
    WarpScratch* scratch;
};
__device__ __forceinline__ SmemLayout carve_smem(unsigned char* raw, bool with_verts) {
    SmemLayout L;
    L.verts = reinterpret_cast<float4*>(raw);
    size_t off = with_verts ? ((size_t)c_sc.n_verts * sizeof(float4) + 15) & ~(size_t)15 : 0;
    L.bs = reinterpret_cast<BlockShared*>(raw + off);
    off += (sizeof(BlockShared) + 15) & ~(size_t)15;
    L.scratch = reinterpret_cast<WarpScratch*>(raw + off);
    return L;
}
static size_t smem_bytes_for(int n_verts, int warps) {
    size_t off = ((size_t)n_verts * sizeof(float4) + 15) & ~(size_t)15;
    off += (sizeof(BlockShared) + 15) & ~(size_t)15;
    return off + (size_t)warps * sizeof(WarpScratch);
}
// block prologue: hulls (only for kernels that run GJK) and scene tables -> shared memory
__device__ __forceinline__ SmemLayout block_prologue(unsigned char* raw, bool with_verts) {
    SmemLayout L = carve_smem(raw, with_verts);
    const int tid = threadIdx.x;
    if (with_verts)
        for (int i = tid; i < c_sc.n_verts; i += blockDim.x) L.verts[i] = __ldg(c_sc.verts + i);
    for (int i = tid; i < (int)(sizeof(SceneImage) / 16); i += blockDim.x)
        reinterpret_cast<uint4*>(L.bs)[i] = __ldg(c_sc.scene_img + i);   // BlockShared starts with the SceneImage members
    if (tid < 16) L.bs->stats[tid] = 0.0;
    if (tid < 16) L.bs->counters[tid] = 0ull;
    __syncthreads();
    return L;
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }

// ------------------------------------------------------------------------------------------------------------------
// obstacle poses
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void planet_pose(int o, int index_one, Xf& T) {
    int idx = index_one;
    if (o == 1) {  // coupled planet (ctlp.py:4481-4485)
        idx = (index_one + c_sc.planet_shift) % c_sc.planet_steps;
        if (idx < 0) idx += c_sc.planet_steps;
    }
    float4 p = __ldg(c_sc.planet_pos[o] + idx);
    float4 q = __ldg(c_sc.planet_quat[o] + idx);
    quat_to_mat(q, T.r);
    T.t[0] = p.x; T.t[1] = p.y; T.t[2] = p.z;
}
// ball pose at flight time t from the env's obstacle record (float64 state, float32 geometry; ctlp.py:4109-4116)
__device__ __forceinline__ void ball_pose(const double* ob, double t, Xf& T) {
    T.t[0] = (float)(ob[SM_OB_BALL_P0] + ob[SM_OB_BALL_V0] * t);
    T.t[1] = (float)(ob[SM_OB_BALL_P0 + 1] + ob[SM_OB_BALL_V0 + 1] * t);
    T.t[2] = (float)(ob[SM_OB_BALL_P0 + 2] + ob[SM_OB_BALL_V0 + 2] * t + (0.5 * -9.81) * (t * t));
    euler_to_mat((float)ob[SM_OB_BALL_EULER0], (float)(ob[SM_OB_BALL_EULER0 + 1] + t * ob[SM_OB_BALL_OMEGA]),
                 (float)ob[SM_OB_BALL_EULER0 + 2], T.r);
}

// ------------------------------------------------------------------------------------------------------------------
// minima over pair lists (all 32 lanes call together)
// ------------------------------------------------------------------------------------------------------------------
// Generic capped minimum: pair p in [0, n) is (A(p), B(p)); `query` is the getClosestPoints distance argument
// (points farther away are not returned), `best` the running minimum.  Candidates are visited in order of increasing
// bounding-sphere lower bound, so that the running minimum prunes the rest exactly.
//   mode 0: pairs[p] from a list (static / self pairs; ctlp.py:3294-3374)
//   mode 1: p -> (rshapes[p / cnt_o], off_o + p % cnt_o) (moving obstacle o; ctlp.py:3258-3280)
__device__ __noinline__ float min_pairs(const float4* verts, const SceneSmem& sm, int mode, const short (*pairs)[2],
                                        const short* rshapes, int n, int cnt_o, int off_o, float query, float best,
                                        const Xf* robot, const Xf* obst, int lane, GjkCounters* cnt) {
#pragma unroll 1
    for (int base = 0; base < n; base += 32) {
        int p = base + lane;
        float lb = FLT_MAX;
        int ia = 0, ib = 0;
        if (p < n) {
            if (mode == 0) { ia = pairs[p][0]; ib = pairs[p][1]; }
            else { ia = rshapes[p / cnt_o]; ib = off_o + p % cnt_o; }
            lb = pair_lower_bound(sm, ia, ib, robot, obst);
        }
#pragma unroll 1
        while (true) {
            float lim = fminf(best, query);
            float cand = lb < lim ? lb : FLT_MAX;
            unsigned key = fkey(-cand);
            unsigned mx = __reduce_max_sync(FULL, key);
            if (mx == fkey(-FLT_MAX)) break;
            int src = __ffs(__ballot_sync(FULL, key == mx)) - 1;
            int qa = __shfl_sync(FULL, ia, src), qb = __shfl_sync(FULL, ib, src);
            float d = pair_distance(verts, sm, qa, qb, robot, obst, lim, -1.f, lane, cnt);
            if (d <= query && d < best) best = d;
            if (lane == src) lb = FLT_MAX;
            if (mode == 1 && best <= 0.f) return 0.f;  // ctlp.py:3277-3278
        }
    }
    return best;
}

// true if some robot shape is within the manifold contact threshold of obstacle o (ctlp.py:4570-4579, :4186-4194;
// SURVEY Appendix B.5)
__device__ __noinline__ bool contact_exists(const float4* verts, const SceneSmem& sm, int o, const Xf* robot,
                                            const Xf* obst, int lane, GjkCounters* cnt) {
    int cnt_o = c_sc.obst_shape_cnt[o], off_o = c_sc.obst_shape_off[o];
    int n = c_sc.n_mov_contact * cnt_o;
#pragma unroll 1
    for (int base = 0; base < n; base += 32) {
        int p = base + lane;
        bool cand = false;
        int ia = 0, ib = 0;
        float th = 0.f;
        if (p < n) {
            int slot = p / cnt_o;
            ia = sm.mov_contact[slot];
            ib = off_o + p % cnt_o;
            th = sm.contact_thresh[o][slot];
            cand = pair_lower_bound(sm, ia, ib, robot, obst) <= th;
        }
        unsigned mask = __ballot_sync(FULL, cand);
#pragma unroll 1
        while (mask) {
            int src = __ffs(mask) - 1;
            mask &= mask - 1;
            int qa = __shfl_sync(FULL, ia, src), qb = __shfl_sync(FULL, ib, src);
            float t = __shfl_sync(FULL, th, src);
            float d = pair_distance(verts, sm, qa, qb, robot, obst, t + 1e-3f, t, lane, cnt);
            if (d <= t) return true;
        }
    }
    return false;
}

// static, self and moving-obstacle distances of the pose whose frames are in `robot` / `obst`
// (get_minimum_distance ctlp.py:3282-3374, get_minimum_distance_to_moving_obstacles :3217-3256)
__device__ __forceinline__ void all_distances(const float4* verts, const SceneSmem& sm, const Xf* robot, const Xf* obst,
                                              bool latched, bool ball_inactive, float& d_static, float& d_self,
                                              float& d_moving, int lane, GjkCounters* cnt) {
    const float cap = (float)c_sc.static_cap, query = (float)c_sc.moving_query;
    d_static = min_pairs(verts, sm, 0, sm.static_pairs, nullptr, c_sc.n_static_pairs, 1, 0, cap, cap, robot, obst, lane, cnt);
    d_self = min_pairs(verts, sm, 0, sm.self_pairs, nullptr, c_sc.n_self_pairs, 1, 0, cap, cap, robot, obst, lane, cnt);
    d_moving = query + 0.002f;  // ctlp.py:3259-3261
    if (latched) {
        d_moving = 0.0f;  // ctlp.py:3224-3234
    } else if (c_sc.n_mov_reward > 0 && !ball_inactive) {
#pragma unroll 1
        for (int o = 0; o < c_sc.n_obstacles; ++o) {
            d_moving = min_pairs(verts, sm, 1, nullptr, sm.mov_reward, c_sc.n_mov_reward * c_sc.obst_shape_cnt[o],
                                 c_sc.obst_shape_cnt[o], c_sc.obst_shape_off[o], query, d_moving, robot, obst, lane, cnt);
            if (d_moving <= 0.0f) break;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// observation (observations.py:233-351, ctlp.py:2352-2391): lane i writes entry i
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float clip1(double x) {
    float f = (float)x;
    return fminf(1.0f, fmaxf(-1.0f, f));
}
__device__ __forceinline__ double normalize_mm(double x, double lo, double hi) {
    return xadd(-1.0, xdiv(xmul(2.0, xsub(x, lo)), xsub(hi, lo)));
}
// kin: the env's kinematic record in global memory (q[8] v[8] a[8] ...); ob: obstacle record (16 doubles)
// tp: the env's target-point record (NULL unless the scene uses target points)
// The entries are spread over `stride` lanes, `lane` = 0 .. stride - 1 (a warp per env, or 8 lanes per env).
// hobs: the nested env's observation (Human scene): its kinematic part is the human's share of this observation
// (Human.kinematic_observation, observations.py:100-110, :294-307)
__device__ __noinline__ void write_observation(float* obs, const double* kin, const double* ob, const double* tp,
                                               int lane, int stride = 32, const float* hobs = nullptr) {
    const int nj = c_sc.n_joints;
    const int n_tp = (c_sc.use_target_points && tp) ? 3 * c_sc.obs_add_tp_pos + 3 * c_sc.obs_add_tp_rel : 0;
#pragma unroll 1
    for (int i = lane; i < c_sc.obs_size; i += stride) {
        double val = 0.0;
        if (i < 3 * nj) {
            int grp = i / nj, j = i - grp * nj;
            double x = kin[grp * 8 + j];
            val = grp == 0 ? normalize_mm(x, c_sc.pos_lo[j], c_sc.pos_hi[j])
                           : xdiv(x, grp == 1 ? c_sc.vel_max[j] : c_sc.acc_max[j]);
        } else if (i < 3 * nj + n_tp) {  // observations.py:326-340; ctlp.py:2247-2271
            int r = i - 3 * nj;
            if (c_sc.obs_add_tp_pos && r < 3) {
                val = normalize_mm(tp[SM_TP_POS + r], c_sc.tp_box_min[r], c_sc.tp_box_max[r]);
            } else {
                if (c_sc.obs_add_tp_pos) r -= 3;
                val = normalize_mm(xsub(tp[SM_TP_POS + r], tp[SM_TP_LINK_POS + r]), c_sc.tp_rel_min[r], c_sc.tp_rel_max[r]);
            }
        } else {
            int r = i - 3 * nj - n_tp;
            if (c_sc.hu.enabled) {
                if (hobs) obs[i] = hobs[r];
                continue;
            }
            if (c_sc.n_obstacles > 0 && c_sc.obst_kind[0] == SM_OBST_BALL) {
                double t = ob[SM_OB_BALL_T];
                if (r < 3) {
                    double p = xadd(ob[SM_OB_BALL_P0 + r], xmul(ob[SM_OB_BALL_V0 + r], t));
                    if (r == 2) p = xadd(p, xmul(0.5 * -9.81, xmul(t, t)));
                    val = normalize_mm(p, c_sc.ball_obs_pos_min[r], c_sc.ball_obs_pos_max[r]);
                } else {
                    int c = r - 3;
                    double vel = xadd(ob[SM_OB_BALL_V0 + c], xmul(c == 2 ? -9.81 : 0.0, t));
                    val = normalize_mm(vel, c_sc.ball_obs_vel_min[c], c_sc.ball_obs_vel_max[c]);
                }
            } else if (c_sc.n_obstacles > 0 && c_sc.obst_kind[0] == SM_OBST_PLANET) {
                int idx = (int)ob[SM_OB_INDEX];
                if (c_sc.obs_planet_size == 1) {
                    val = normalize_mm((double)idx, 0.0, (double)c_sc.planet_steps);
                } else {
                    double half = c_sc.planet_obs_half[r];
                    val = normalize_mm(c_sc.planet_local_xy[2 * idx + r], -half, half);
                }
            }
        }
        obs[i] = clip1(val);
    }
}

// target link point of the pose q (float64 joint angles in global memory), computed serially by one thread from the
// constant-memory chain (LinkPointBase.get_position, ctlp.py:4962-5075); used where an episode starts
__device__ __noinline__ V3 target_link_point_serial(const double* q) {
    Xf F;
    xf_identity(F);
#pragma unroll 1
    for (int j = 0; j < c_sc.n_joints; ++j) {
        float s, c;
        sincosf((float)q[j], &s, &c);
        Xf L, C;
        float Rj[9];
        axis_angle(c_sc.jaxis[j][0], c_sc.jaxis[j][1], c_sc.jaxis[j][2], c, s, Rj);
        const float* A = c_sc.jR[j];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k)
                L.r[3 * i + k] = fmaf(A[3 * i], Rj[k], fmaf(A[3 * i + 1], Rj[3 + k], A[3 * i + 2] * Rj[6 + k]));
        L.t[0] = c_sc.jt[j][0]; L.t[1] = c_sc.jt[j][1]; L.t[2] = c_sc.jt[j][2];
        xf_compose(F, L, C);
        F = C;
    }
    return xf_apply(F, c_sc.tp_local[0], c_sc.tp_local[1], c_sc.tp_local[2]);
}

// start of an episode with target points: link point of the start pose, first target point (given, or drawn from the
// pool with the env's Philox counter), distances (ctlp.py:2216-2245).  One thread.
__device__ __forceinline__ void target_episode_start(double* tp, const double* kin, const double* first_target,
                                                     const double* pool, int pool_n, int env, uint32_t k0, uint32_t k1) {
    const V3 p = target_link_point_serial(kin);
    double draws = tp[SM_TP_DRAWS];
    double t[3] = {0.0, 0.0, 0.0};
    if (first_target) { t[0] = first_target[0]; t[1] = first_target[1]; t[2] = first_target[2]; }
    else if (pool && pool_n > 0) {
        uint4 r = philox((uint32_t)env, (uint32_t)draws, 0x7A26u, 2u, k0, k1);
        const double* e = pool + (size_t)(r.x % (uint32_t)pool_n) * 4;
        t[0] = e[0]; t[1] = e[1]; t[2] = e[2];
    }
    const double dx = t[0] - (double)p.x, dy = t[1] - (double)p.y, dz = t[2] - (double)p.z;
    const double dist = sqrt(dx * dx + dy * dy + dz * dz);
    tp[SM_TP_POS] = t[0]; tp[SM_TP_POS + 1] = t[1]; tp[SM_TP_POS + 2] = t[2];
    tp[SM_TP_LAST_DIST] = dist; tp[SM_TP_INIT_DIST] = dist; tp[SM_TP_ACTIVE] = 1.0; tp[SM_TP_REACHED_N] = 0.0;
    tp[SM_TP_LINK_POS] = (double)p.x; tp[SM_TP_LINK_POS + 1] = (double)p.y; tp[SM_TP_LINK_POS + 2] = (double)p.z;
    tp[SM_TP_DRAWS] = draws + 1.0; tp[SM_TP_REACHED] = 0.0;
}

// lane j holds joint angle q_j in float64 (the state): sin/cos in float64, chain in float32
__device__ __forceinline__ void frames_from_q64(const SceneSmem& sm, double q, Xf* out, int lane) {
    double s, c;
    sincos(q, &s, &c);
    fk_scan(sm, (float)c, (float)s, out, lane);
}
__device__ __forceinline__ void frames_from_q32f(const SceneSmem& sm, float q, Xf* out, int lane) {
    float s, c;
    sincosf(q, &s, &c);
    fk_scan(sm, c, s, out, lane);
}
__device__ __forceinline__ void frames_from_q32(const SceneSmem& sm, const float* qrow, Xf* out, int lane) {
    float s, c;
    sincosf(qrow[lane & 7], &s, &c);
    fk_scan(sm, c, s, out, lane);
}

// ------------------------------------------------------------------------------------------------------------------
// forward kinematics of the human: frame 0 = base, 1 + j = child link of joint j; two serial arms of four joints
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void human_base(Xf& B) {
#pragma unroll
    for (int i = 0; i < 9; ++i) B.r[i] = c_sc.hu.baseR[i];
    B.t[0] = c_sc.hu.baset[0]; B.t[1] = c_sc.hu.baset[1]; B.t[2] = c_sc.hu.baset[2];
}
// one joint of an arm: F <- F o [R_fix Rot(axis_j, q) | t_fix]; j is uniform over the calling lanes' loop
__device__ __forceinline__ void human_chain_step(Xf& F, int j, float q) {
    float s, c;
    sincosf(q, &s, &c);
    Xf L, C;
    float Rj[9];
    axis_angle(c_sc.hu.jaxis[j][0], c_sc.hu.jaxis[j][1], c_sc.hu.jaxis[j][2], c, s, Rj);
    const float* A = c_sc.hu.jR[j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k)
            L.r[3 * i + k] = fmaf(A[3 * i], Rj[k], fmaf(A[3 * i + 1], Rj[3 + k], A[3 * i + 2] * Rj[6 + k]));
    L.t[0] = c_sc.hu.jt[j][0]; L.t[1] = c_sc.hu.jt[j][1]; L.t[2] = c_sc.hu.jt[j][2];
    xf_compose(F, L, C);
    F = C;
}
// all nine frames by one thread (q: 8 joint angles)
__device__ __forceinline__ void human_fk_serial(const float* q, Xf* fr) {
    human_base(fr[0]);
#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
        Xf F = fr[0];
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            human_chain_step(F, 4 * r + i, q[4 * r + i]);
            fr[1 + 4 * r + i] = F;
        }
    }
}
// warp version: lanes 0..7 each hold their joint's angle; the two arms are scanned as segments of four lanes
__device__ __noinline__ void human_fk_scan(float q, Xf* out, int lane) {
    const int j = lane & 7;
    float s, c;
    sincosf(q, &s, &c);
    Xf X;
    {
        float Rj[9];
        axis_angle(c_sc.hu.jaxis[j][0], c_sc.hu.jaxis[j][1], c_sc.hu.jaxis[j][2], c, s, Rj);
        const float* A = c_sc.hu.jR[j];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k)
                X.r[3 * i + k] = fmaf(A[3 * i], Rj[k], fmaf(A[3 * i + 1], Rj[3 + k], A[3 * i + 2] * Rj[6 + k]));
        X.t[0] = c_sc.hu.jt[j][0]; X.t[1] = c_sc.hu.jt[j][1]; X.t[2] = c_sc.hu.jt[j][2];
    }
#pragma unroll
    for (int d = 1; d < 4; d <<= 1) {
        Xf P, C;
#pragma unroll
        for (int i = 0; i < 9; ++i) P.r[i] = __shfl_up_sync(FULL, X.r[i], d, 4);
#pragma unroll
        for (int i = 0; i < 3; ++i) P.t[i] = __shfl_up_sync(FULL, X.t[i], d, 4);
        xf_compose(P, X, C);
        if ((j & 3) >= d) X = C;
    }
    Xf B, W;
    human_base(B);
    xf_compose(B, X, W);
    if (lane == 0) out[0] = B;
    if (lane < SM_HUMAN_JOINTS) out[1 + lane] = W;
    __syncwarp();
}
__device__ __forceinline__ V3 human_link_point(const Xf* fr, int r) {
    return xf_apply(fr[4 * (r + 1)], c_sc.hu.tp_local[r][0], c_sc.hu.tp_local[r][1], c_sc.hu.tp_local[r][2]);
}

