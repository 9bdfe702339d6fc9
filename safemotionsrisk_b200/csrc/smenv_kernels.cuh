// smenv_kernels.cuh -- the env-step kernel and its helpers (one warp per environment).
#pragma once
#include "smenv_device.cuh"

#define SM_WARPS_PER_BLOCK 8
#define SM_POOL_STRIDE 48  /* doubles per start-pool entry: kin record (32) + obstacle record (16) */
#define SM_BALL_STRIDE 12  /* doubles per ball-pool entry: p0 v0 euler0 omega nmax nhit */
#define SM_MAX_SUB 32

// per-warp scratch in shared memory
struct WarpScratch {
    float qsub[SM_MAX_SUB][SM_MAX_JOINTS];  // tracked joint pose seen by the contact test of each sub-step
    Xf fr[1 + SM_MAX_JOINTS];               // robot frames at the end-of-step setpoint pose
    Xf fr2[1 + SM_MAX_JOINTS];              // robot frames of one sub-step (narrow phase)
    Xf ob[SM_MAX_OBSTACLES];                // obstacle poses at the end of the step
    Xf ob2[SM_MAX_OBSTACLES];               // obstacle poses of one sub-step
};

struct BlockShared {
    double stats[16];
    unsigned long long counters[6];
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }

// ------------------------------------------------------------------------------------------------------------------
// obstacle poses
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int planet_index(int o, int index_one) {
    int idx = index_one;
    if (o == 1) {  // ctlp.py:4481-4485
        idx = (index_one + c_sc.planet_shift) % c_sc.planet_steps;
        if (idx < 0) idx += c_sc.planet_steps;
    }
    return idx;
}
__device__ __forceinline__ void planet_pose(int o, int index_one, Xf& T) {
    int idx = planet_index(o, index_one);
    float4 p = __ldg(c_sc.planet_pos[o] + idx);
    float4 q = __ldg(c_sc.planet_quat[o] + idx);
    quat_to_mat(q, T.r);
    T.t[0] = p.x; T.t[1] = p.y; T.t[2] = p.z;
}
// ball record values are passed in float64 (bit-exact state), the pose is float32 geometry
__device__ __forceinline__ void ball_pose(const double* p0, const double* v0, const double* e0, double omega, double t,
                                          Xf& T) {
    T.t[0] = (float)(p0[0] + v0[0] * t);
    T.t[1] = (float)(p0[1] + v0[1] * t);
    T.t[2] = (float)(p0[2] + v0[2] * t + (0.5 * -9.81) * (t * t));
    euler_to_mat((float)e0[0], (float)(e0[1] + t * omega), (float)e0[2], T.r);  // ctlp.py:4112-4114
}

// ------------------------------------------------------------------------------------------------------------------
// pair distance helpers (all 32 lanes of the warp call these together)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ const Xf& frame_of(const DevShape& sh, const Xf* robot, const Xf* obst) {
    return sh.frame >= 100 ? obst[sh.frame - 100] : robot[sh.frame];
}

// distance between shapes ia and ib minus both margins; `upper` / `touch` are in the same (margin-free) metric
__device__ __forceinline__ float pair_distance(const float4* verts, int ia, int ib, const Xf* robot, const Xf* obst,
                                               float upper, float touch, int lane, GjkCounters* cnt) {
    const DevShape& A = c_sc.shapes[ia];
    const DevShape& B = c_sc.shapes[ib];
    Xf TA = frame_of(A, robot, obst);
    Xf TB = frame_of(B, robot, obst);
    float m = A.margin + B.margin;
    V3 ca = xf_apply(TA, A.cx, A.cy, A.cz), cb = xf_apply(TB, B.cx, B.cy, B.cz);
    float d = gjk_warp(verts + A.off, A.cnt, TA, verts + B.off, B.cnt, TB, ca - cb, upper > 0.f ? upper + m : 0.f,
                       touch >= 0.f ? touch + m : -1.f, lane, cnt);
    return d - m;
}

// bounding-sphere lower bound of the same quantity
__device__ __forceinline__ float pair_lower_bound(int ia, int ib, const Xf* robot, const Xf* obst) {
    const DevShape& A = c_sc.shapes[ia];
    const DevShape& B = c_sc.shapes[ib];
    const Xf& TA = frame_of(A, robot, obst);
    const Xf& TB = frame_of(B, robot, obst);
    V3 ca = xf_apply(TA, A.cx, A.cy, A.cz);
    if (B.frame == 0) {  // static shape in the world frame: sphere against its axis-aligned box (tight for the table)
        float dx = fmaxf(fmaxf(B.bmin[0] - ca.x, ca.x - B.bmax[0]), 0.f);
        float dy = fmaxf(fmaxf(B.bmin[1] - ca.y, ca.y - B.bmax[1]), 0.f);
        float dz = fmaxf(fmaxf(B.bmin[2] - ca.z, ca.z - B.bmax[2]), 0.f);
        return sqrtf(dx * dx + dy * dy + dz * dz) - A.radius - A.margin - B.margin;
    }
    V3 d = ca - xf_apply(TB, B.cx, B.cy, B.cz);
    return sqrtf(dot(d, d)) - A.radius - B.radius - A.margin - B.margin;
}

// min over a pair list, capped (get_minimum_distance, ctlp.py:3282-3374): start at cap, only d <= cap counts
__device__ float min_pair_list(const float4* verts, const short (*pairs)[2], int n, float cap, const Xf* robot,
                               const Xf* obst, int lane, GjkCounters* cnt, unsigned* culled) {
    float best = cap;
    for (int base = 0; base < n; base += 32) {
        int p = base + lane;
        float lb = FLT_MAX;
        if (p < n) lb = pair_lower_bound(pairs[p][0], pairs[p][1], robot, obst);
        // visit candidates in order of increasing lower bound so that the running minimum prunes the rest
        while (true) {
            float cand = lb < best ? lb : FLT_MAX;
            unsigned key = fkey(-cand);
            unsigned mx = __reduce_max_sync(FULL, key);
            if (mx == fkey(-FLT_MAX)) break;
            int src = __ffs(__ballot_sync(FULL, key == mx)) - 1;
            int q = base + src;
            float d = pair_distance(verts, pairs[q][0], pairs[q][1], robot, obst, best, -1.f, lane, cnt);
            if (d <= cap && d < best) best = d;
            if (lane == src) lb = FLT_MAX;
        }
        if (culled && p < n && lb != FLT_MAX) atomicAdd(culled, 1u);
    }
    return best;
}

// min distance between the robot shapes in `rshapes` and every part of obstacle o (ctlp.py:3258-3280)
__device__ float min_moving(const float4* verts, const short* rshapes, int nr, int o, float query, float best,
                            const Xf* robot, const Xf* obst, int lane, GjkCounters* cnt, unsigned* culled) {
    int cnt_o = c_sc.obst_shape_cnt[o], off_o = c_sc.obst_shape_off[o];
    int n = nr * cnt_o;
    // whole-obstacle cull first
    for (int base = 0; base < n; base += 32) {
        int p = base + lane;
        float lb = FLT_MAX;
        int ia = 0, ib = 0;
        if (p < n) {
            ia = rshapes[p / cnt_o];
            ib = off_o + p % cnt_o;
            lb = pair_lower_bound(ia, ib, robot, obst);
        }
        while (true) {
            float lim = fminf(best, query);
            float cand = lb < lim ? lb : FLT_MAX;
            unsigned key = fkey(-cand);
            unsigned mx = __reduce_max_sync(FULL, key);
            if (mx == fkey(-FLT_MAX)) break;
            int src = __ffs(__ballot_sync(FULL, key == mx)) - 1;
            int qa = __shfl_sync(FULL, ia, src), qb = __shfl_sync(FULL, ib, src);
            float d = pair_distance(verts, qa, qb, robot, obst, lim, -1.f, lane, cnt);
            if (d <= query && d < best) best = d;
            if (lane == src) lb = FLT_MAX;
            if (best <= 0.f) return 0.f;  // ctlp.py:3277-3278
        }
        if (culled && p < n && lb != FLT_MAX) atomicAdd(culled, 1u);
    }
    return best;
}

// true if some robot shape is within the manifold contact threshold of obstacle o (ctlp.py:4570-4579, :4186-4194)
__device__ bool contact_exists(const float4* verts, int o, const Xf* robot, const Xf* obst, int lane,
                               GjkCounters* cnt) {
    int cnt_o = c_sc.obst_shape_cnt[o], off_o = c_sc.obst_shape_off[o];
    int n = c_sc.n_mov_contact * cnt_o;
    for (int base = 0; base < n; base += 32) {
        int p = base + lane;
        bool cand = false;
        int ia = 0, ib = 0;
        float th = 0.f;
        if (p < n) {
            int slot = p / cnt_o;
            ia = c_sc.mov_contact[slot];
            ib = off_o + p % cnt_o;
            th = c_sc.contact_thresh[o][slot];
            cand = pair_lower_bound(ia, ib, robot, obst) <= th;
        }
        unsigned mask = __ballot_sync(FULL, cand);
        while (mask) {
            int src = __ffs(mask) - 1;
            mask &= mask - 1;
            int qa = __shfl_sync(FULL, ia, src), qb = __shfl_sync(FULL, ib, src);
            float t = __shfl_sync(FULL, th, src);
            float d = pair_distance(verts, qa, qb, robot, obst, t + 1e-3f, t, lane, cnt);
            if (d <= t) return true;
        }
    }
    return false;
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
// q, v, a: value of joint (lane & 7) on every lane; ob: obstacle slot `lane` on lanes 0..15
__device__ void write_observation(float* obs, double q, double v, double a, double ob, int lane) {
    int nj = c_sc.n_joints;
    // every lane first gathers the ball / planet scalars (shuffles must be warp-uniform)
    double p0x = shfl_d(ob, SM_OB_BALL_P0), p0y = shfl_d(ob, SM_OB_BALL_P0 + 1), p0z = shfl_d(ob, SM_OB_BALL_P0 + 2);
    double v0x = shfl_d(ob, SM_OB_BALL_V0), v0y = shfl_d(ob, SM_OB_BALL_V0 + 1), v0z = shfl_d(ob, SM_OB_BALL_V0 + 2);
    double t = shfl_d(ob, SM_OB_BALL_T);
    int idx = (int)shfl_d(ob, SM_OB_INDEX);
    for (int base = 0; base < c_sc.obs_size; base += 32) {
        int i = base + lane;
        int grp = i / nj, j = i - grp * nj;
        // kinematic part: lane i needs joint j of group grp; all lanes hold joint (lane & 7)
        double qj = shfl_d(q, j & 7), vj = shfl_d(v, j & 7), aj = shfl_d(a, j & 7);
        if (i >= c_sc.obs_size) continue;
        double val = 0.0;
        if (i < nj) val = normalize_mm(qj, c_sc.pos_lo[j], c_sc.pos_hi[j]);
        else if (i < 2 * nj) val = xdiv(vj, c_sc.vel_max[j]);
        else if (i < 3 * nj) val = xdiv(aj, c_sc.acc_max[j]);
        else {
            int r = i - 3 * nj;
            if (c_sc.n_obstacles > 0 && c_sc.obst_kind[0] == SM_OBST_BALL) {
                if (r < 3) {
                    double p = r == 0 ? xadd(p0x, xmul(v0x, t)) : r == 1 ? xadd(p0y, xmul(v0y, t))
                        : xadd(xadd(p0z, xmul(v0z, t)), xmul(0.5 * -9.81, xmul(t, t)));
                    val = normalize_mm(p, c_sc.ball_obs_pos_min[r], c_sc.ball_obs_pos_max[r]);
                } else {
                    int c = r - 3;
                    double vel = c == 0 ? xadd(v0x, xmul(0.0, t)) : c == 1 ? xadd(v0y, xmul(0.0, t))
                        : xadd(v0z, xmul(-9.81, t));
                    val = normalize_mm(vel, c_sc.ball_obs_vel_min[c], c_sc.ball_obs_vel_max[c]);
                }
            } else if (c_sc.n_obstacles > 0 && c_sc.obst_kind[0] == SM_OBST_PLANET) {
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

// ------------------------------------------------------------------------------------------------------------------
// forward kinematics into shared memory (all lanes run the same chain; lane j stores frame j)
// T_frame = T_parent * [R_fix | t_fix] * Rot(axis, q)  (ctlp.py:2940-2988, LinkBase.get_position :5163-5195)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fk_to_smem(const float* cq, const float* sq, Xf* out, int lane) {
    if (lane == 0) {
        Xf w;
#pragma unroll
        for (int i = 0; i < 9; ++i) w.r[i] = (i % 4 == 0) ? 1.0f : 0.0f;
        w.t[0] = w.t[1] = w.t[2] = 0.0f;
        out[0] = w;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < SM_MAX_JOINTS; ++j) {
        if (j < c_sc.n_joints) {
            Xf P = out[c_sc.joint_parent[j]];
            float R1[9], Rj[9];
            Xf F;
            mat_mul(P.r, c_sc.jR[j], R1);
            V3 tp = xf_apply(P, c_sc.jt[j][0], c_sc.jt[j][1], c_sc.jt[j][2]);
            axis_angle(c_sc.jaxis[j], cq[j], sq[j], Rj);
            mat_mul(R1, Rj, F.r);
            F.t[0] = tp.x; F.t[1] = tp.y; F.t[2] = tp.z;
            if (lane == j) out[1 + j] = F;
            __syncwarp();
        }
    }
}
// lane j holds joint angle q_j in float64 (the state); sin/cos in float64, chain in float32
__device__ __forceinline__ void frames_from_q64(double q, Xf* out, int lane) {
    double s, c;
    sincos(q, &s, &c);
    float cf = (float)c, sf = (float)s;
    float cq[SM_MAX_JOINTS], sq[SM_MAX_JOINTS];
#pragma unroll
    for (int j = 0; j < SM_MAX_JOINTS; ++j) {
        cq[j] = __shfl_sync(FULL, cf, j);
        sq[j] = __shfl_sync(FULL, sf, j);
    }
    fk_to_smem(cq, sq, out, lane);
}
__device__ __forceinline__ void frames_from_q32(const float* qrow, Xf* out, int lane) {
    float ql = qrow[lane & 7], sl, cl;
    sincosf(ql, &sl, &cl);
    float cq[SM_MAX_JOINTS], sq[SM_MAX_JOINTS];
#pragma unroll
    for (int j = 0; j < SM_MAX_JOINTS; ++j) {
        cq[j] = __shfl_sync(FULL, cl, j);
        sq[j] = __shfl_sync(FULL, sl, j);
    }
    fk_to_smem(cq, sq, out, lane);
}
