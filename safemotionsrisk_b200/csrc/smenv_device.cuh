// smenv_device.cuh -- device-side building blocks of the env-step kernels (sm_100a).
//
// Layout of the work: ONE WARP PER ENVIRONMENT.
//   * float64 joint-space maths (safe range, action mapping, interpolation, motor tracking) runs on lanes 0..n_joints-1
//     with explicit round-to-nearest intrinsics (no FMA contraction), so the joint trajectory is bit-identical to a
//     strict IEEE evaluation on the host;
//   * float32 geometry (forward kinematics, GJK) is warp-cooperative: the 32 lanes split the hull vertices of a
//     support search (vertices live in shared memory as float4) and reduce with one REDUX + ballot;
//   * the 24 physics sub-steps of one env step are mapped to lanes for the broad phase (lane k = sub-step k).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "../../include/smenv.h"

#define FULL 0xffffffffu
#define SM_BIG 1.0e6

// ------------------------------------------------------------------------------------------------------------------
// scene constants on the device
// ------------------------------------------------------------------------------------------------------------------
struct DevShape {
    int frame, off, cnt, link;
    float margin, cx, cy, cz, radius;
    float bmin[3], bmax[3];  // axis-aligned box of the core vertices in frame coordinates
};

struct DevScene {
    int n_joints, substeps, contact_stride, limit_velocity, limit_position;
    int joint_parent[SM_MAX_JOINTS];
    float jR[SM_MAX_JOINTS][9], jt[SM_MAX_JOINTS][3], jaxis[SM_MAX_JOINTS][3];
    double pos_lo[SM_MAX_JOINTS], pos_hi[SM_MAX_JOINTS], vel_max[SM_MAX_JOINTS], acc_max[SM_MAX_JOINTS],
        jerk_max[SM_MAX_JOINTS];
    double ts, action_mapping_factor, track_kp, track_vel;
    int n_shapes, n_verts;
    DevShape shapes[SM_MAX_SHAPES];
    int n_static_pairs, n_self_pairs, n_mov_reward, n_mov_contact;
    int contact_frame_start[SM_MAX_JOINTS + 2];  // mov_contact slots sorted by frame: slots of frame f
    short static_pairs[SM_MAX_PAIRS][2], self_pairs[SM_MAX_PAIRS][2];
    short mov_reward[SM_MAX_MOV_ROBOT], mov_contact[SM_MAX_MOV_ROBOT];
    int n_obstacles;
    int obst_kind[SM_MAX_OBSTACLES], obst_shape_off[SM_MAX_OBSTACLES], obst_shape_cnt[SM_MAX_OBSTACLES];
    float obst_center[SM_MAX_OBSTACLES][3], obst_radius[SM_MAX_OBSTACLES];
    float contact_thresh[SM_MAX_OBSTACLES][SM_MAX_MOV_ROBOT];
    float contact_thresh_max[SM_MAX_OBSTACLES];
    int planet_steps, planet_shift, obs_planet_size;
    const float4* planet_pos[SM_MAX_OBSTACLES];   // xyz, w unused
    const float4* planet_quat[SM_MAX_OBSTACLES];  // xyzw
    const double* planet_local_xy;
    double planet_obs_half[2];
    double ball_obs_pos_min[3], ball_obs_pos_max[3], ball_obs_vel_min[3], ball_obs_vel_max[3];
    double ball_active_xy;
    double static_cap, moving_query, collision_dist;
    double w_self, w_static, w_moving, d_self, d_static, d_moving, w_low_acc, thr_low_acc, w_low_vel, thr_low_vel;
    int punish_action, terminate_self, terminate_static, terminate_moving;
    double action_thresh, action_max_punishment, termination_bonus, early_termination_punishment;
    int episode_steps, obs_size;
    // sampling
    double start_box_min[3], start_box_max[3];
    float target_offset[3], target_R[9], target_t[3];
    double kinematic_sampling_probability, stay_in_state_probability, min_start_distance;
    double ball_sphere_center[3], ball_sphere_radius, ball_height_min, ball_height_max, ball_angle_min, ball_angle_max;
    double ball_speed, ball_radius, ball_high_angle_probability;
    double ball_target_box_min[3], ball_target_box_max[3], ball_invalid_min[3], ball_invalid_max[3];
    double ball_final_min[3], ball_final_max[3], plane_z;
    int ball_check_invalid, ball_random_initial, has_table;
    double min_start_self, ball_target_min_static, ball_target_min_self;
    const float4* verts;  // device, n_verts
};

// the library is one translation unit (smenv.cu), so the constant-memory scene is defined here
__constant__ DevScene c_sc;

// ------------------------------------------------------------------------------------------------------------------
// strict IEEE float64 helpers: never contracted into FMA, same rounding as the host evaluating one operation at a time
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double xsqrt(double a) { return __dsqrt_rn(a); }

// ------------------------------------------------------------------------------------------------------------------
// 1. safe acceleration range (actions.py:97-106, :206-211 -> klimits.PosVelJerkLimitation; model in DESIGN.md)
// ------------------------------------------------------------------------------------------------------------------
enum { CODE_VEL_HI = 1, CODE_VEL_LO = 2, CODE_POS_HI = 4, CODE_POS_LO = 8, CODE_ACC = 16 };

// largest next-knot acceleration that keeps the velocity peak under vmax when braking as hard as allowed afterwards
__device__ __forceinline__ double vel_upper(double v0, double a0, double vmax, double J, double A, double ts) {
    double c = xsub(xadd(v0, xmul(xmul(a0, ts), 0.5)), vmax);
    if (c > 0.0) {
        double den = xsub(vmax, v0);
        if (a0 <= 0.0 || den <= 0.0) return -SM_BIG;
        return xsub(a0, xdiv(xmul(xmul(a0, a0), ts), xmul(2.0, den)));
    }
    double a1u = xmul(J, xsub(xsqrt(xsub(xmul(xmul(ts, ts), 0.25), xdiv(xmul(2.0, c), J))), xmul(ts, 0.5)));
    double JT = xmul(J, ts);
    double delta = xsub(JT, A);
    if (delta <= 0.0) return a1u;
    double n = xsub(ceil(xdiv(a1u, JT)), 1.0);
    if (n < 0.0) n = 0.0;
    double x = xsub(a1u, xmul(n, JT));
    if (x >= delta) return a1u;
    double C = xadd(c, xmul(xmul(xmul(xmul(JT, ts), n), xadd(n, 1.0)), 0.5));
    double D = xmul(ts, xadd(n, 0.5));
    double qa = xadd(D, xmul(ts, 0.5)), qb = xadd(C, xmul(D, A)), qc = xmul(C, A);
    x = xdiv(xsub(xsqrt(xsub(xmul(qb, qb), xmul(xmul(4.0, qa), qc))), qb), xmul(2.0, qa));
    return xadd(xmul(n, JT), x);
}

// highest position reached when the next knot acceleration is a1 and the hardest admissible braking follows
__device__ double pos_peak(double p, double v, double a, double a1, double J, double A, double ts) {
    double best = p;
    double an = a1;
    for (int it = 0; it < 16; ++it) {
        double j = xdiv(xsub(an, a), ts);
        double tau = -1.0;
        if (j == 0.0) {
            if (a < 0.0 && v > 0.0) tau = xdiv(-v, a);
        } else {
            double disc = xsub(xmul(a, a), xmul(xmul(2.0, j), v));
            if (disc >= 0.0) {
                double s = xsqrt(disc);
                if (a <= 0.0) {
                    if (xsub(s, a) > 0.0) tau = xdiv(xmul(2.0, v), xsub(s, a));
                } else {
                    tau = xdiv(xsub(-a, s), j);
                }
            }
        }
        if (tau > 0.0 && tau <= ts) {
            // p + v*tau + 0.5*a*tau*tau + (j*tau*tau*tau)/6.0, left to right
            double pk = xadd(xadd(xadd(p, xmul(v, tau)), xmul(xmul(xmul(0.5, a), tau), tau)),
                             xdiv(xmul(xmul(xmul(j, tau), tau), tau), 6.0));
            if (pk > best) best = pk;
        }
        double pn = xadd(xadd(p, xmul(v, ts)), xmul(xmul(xadd(xdiv(a, 3.0), xdiv(an, 6.0)), ts), ts));
        double vn = xadd(v, xmul(xmul(xadd(a, an), ts), 0.5));
        p = pn; v = vn; a = an;
        if (p > best) best = p;
        if (a <= -A) {
            if (v > 0.0) {
                double pk = xadd(p, xdiv(xmul(v, v), xmul(2.0, A)));
                if (pk > best) best = pk;
            }
            break;
        }
        if (v <= 0.0 && a <= 0.0) break;
        an = xsub(a, xmul(J, ts));
        if (an < -A) an = -A;
    }
    return best;
}

__device__ double pos_upper(double p, double v, double a, double pmax, double lo, double hi, double J, double A,
                            double ts) {
    double fr = xsub(pos_peak(p, v, a, hi, J, A, ts), pmax);
    if (fr <= 0.0) return SM_BIG;
    double fl = xsub(pos_peak(p, v, a, lo, J, A, ts), pmax);
    if (fl > 0.0) return fl > 1e-6 ? -SM_BIG : lo;
    double xl = lo, xr = hi;
    int side = 0;
    for (int it = 0; it < 40; ++it) {
        if (xsub(xr, xl) <= 1e-9) break;
        double x = xsub(xr, xdiv(xmul(fr, xsub(xr, xl)), xsub(fr, fl)));
        if (!(x > xl && x < xr)) x = xmul(0.5, xadd(xl, xr));
        double f = xsub(pos_peak(p, v, a, x, J, A, ts), pmax);
        if (f <= 0.0) {
            xl = x; fl = f;
            if (side == -1) fr = xmul(fr, 0.5);
            side = -1;
        } else {
            xr = x; fr = f;
            if (side == 1) fl = xmul(fl, 0.5);
            side = 1;
        }
    }
    return xl;
}

__device__ __forceinline__ void clamp_range(double& lo, double& hi, double blo, double bhi, int code_hi, int code_lo,
                                            int& code) {
    double nhi = hi < bhi ? hi : bhi;
    double nlo = lo > blo ? lo : blo;
    if (nhi < lo) { if (xsub(lo, nhi) > 1e-6) code |= code_hi; nhi = lo; }
    if (nlo > hi) { if (xsub(nlo, hi) > 1e-6) code |= code_lo; nlo = hi; }
    if (nlo > nhi) { if (xsub(nlo, nhi) > 1e-6) code |= code_hi | code_lo; nlo = nhi; }
    lo = nlo; hi = nhi;
}

__device__ void safe_range_joint(int j, double p, double v, double a, double& out_lo, double& out_hi, int& out_code) {
    double ts = c_sc.ts, J = c_sc.jerk_max[j], A = c_sc.acc_max[j], V = c_sc.vel_max[j];
    int code = 0;
    double lo = xsub(a, xmul(J, ts)), hi = xadd(a, xmul(J, ts));
    if (lo < -A) lo = -A;
    if (hi > A) hi = A;
    if (lo > hi) {
        code |= CODE_ACC;
        if (a > 0.0) lo = hi; else hi = lo;
    }
    if (c_sc.limit_velocity) {
        double bhi = vel_upper(v, a, V, J, A, ts);
        double blo = -vel_upper(-v, -a, V, J, A, ts);
        clamp_range(lo, hi, blo, bhi, CODE_VEL_HI, CODE_VEL_LO, code);
    }
    if (c_sc.limit_position) {
        double bhi = pos_upper(p, v, a, c_sc.pos_hi[j], lo, hi, J, A, ts);
        double blo = -pos_upper(-p, -v, -a, -c_sc.pos_lo[j], -hi, -lo, J, A, ts);
        clamp_range(lo, hi, blo, bhi, CODE_POS_HI, CODE_POS_LO, code);
    }
    out_lo = lo; out_hi = hi; out_code = code;
}

// actions.py:268-280
__device__ __forceinline__ double map_action(double u, double lo, double hi) {
    if (c_sc.action_mapping_factor != 1.0) {
        double mf = xmul(0.5, xadd(c_sc.action_mapping_factor, 1.0));
        double diff = xsub(hi, lo);
        hi = xadd(lo, xmul(mf, diff));
        lo = xadd(lo, xmul(xsub(1.0, mf), diff));
    }
    return xadd(lo, xmul(xmul(0.5, xadd(u, 1.0)), xsub(hi, lo)));
}

// np.linspace(ts / S, ts, S)[k-1]  (actions.py:420-421)
__device__ __forceinline__ double substep_time(int k) {
    int S = c_sc.substeps;
    if (S <= 1 || k == S) return c_sc.ts;
    double start = xdiv(c_sc.ts, (double)S);
    double step = xdiv(xsub(c_sc.ts, start), (double)(S - 1));
    return xadd(start, xmul((double)(k - 1), step));
}

// actions.py:468-487
__device__ __forceinline__ void interpolate(double q0, double v0, double a0, double a1, double t, double& q,
                                            double& v, double& a) {
    double jerk = xdiv(xsub(a1, a0), c_sc.ts);
    a = xadd(a0, xmul(jerk, t));
    v = xadd(xadd(v0, xmul(a0, t)), xmul(xmul(xmul(0.5, jerk), t), t));
    q = xadd(xadd(xadd(q0, xmul(v0, t)), xmul(xmul(xmul(0.5, a0), t), t)),
             xmul(xmul(xmul(xmul(1.0 / 6.0, jerk), t), t), t));
}

// ------------------------------------------------------------------------------------------------------------------
// 2. Philox4x32-10 counter-based RNG (Salmon et al., SC'11)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ float u01f(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
__device__ __forceinline__ double u01d(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

// per-warp sampling stream: every draw advances the counter; all lanes get the same numbers
struct Rng {
    uint32_t k0, k1, c0, c1, n;
    __device__ Rng(uint64_t seed, uint32_t a, uint32_t b) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), c0(a), c1(b), n(0) {}
    __device__ uint4 next4() { return philox(c0, c1, n++, 0x5afe, k0, k1); }
    __device__ double uniform() { uint4 r = next4(); return u01d(r.x, r.y); }
    __device__ double uniform(double lo, double hi) { return lo + (hi - lo) * uniform(); }
};

// ------------------------------------------------------------------------------------------------------------------
// 3. float32 rigid transforms and forward kinematics
// ------------------------------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void mat_mul(const float* A, const float* B, float* C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k)
            C[3 * i + k] = fmaf(A[3 * i], B[k], fmaf(A[3 * i + 1], B[3 + k], A[3 * i + 2] * B[6 + k]));
}
__device__ __forceinline__ void axis_angle(const float* ax, float c, float s, float* R) {
    float t = 1.0f - c, x = ax[0], y = ax[1], z = ax[2];
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
__device__ __forceinline__ void euler_to_mat(float e0, float e1, float e2, float* R) {
    float sr, cr, sp, cp, sy, cy;
    sincosf(e0, &sr, &cr); sincosf(e1, &sp, &cp); sincosf(e2, &sy, &cy);
    R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
    R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
    R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
}

// Forward kinematics into `frames` (1 + n_joints entries, entry 0 = world).  cs[j] = (cos q_j, sin q_j).
// T_frame = T_parent * [R_fix | t_fix] * Rot(axis, q)  (ctlp.py:2940-2988, LinkBase.get_position :5163-5195)
__device__ __forceinline__ void fk_frames(const float* cq, const float* sq, Xf* frames) {
    Xf& w = frames[0];
#pragma unroll
    for (int i = 0; i < 9; ++i) w.r[i] = (i % 4 == 0) ? 1.0f : 0.0f;
    w.t[0] = w.t[1] = w.t[2] = 0.0f;
    for (int j = 0; j < c_sc.n_joints; ++j) {
        const Xf& P = frames[c_sc.joint_parent[j]];
        float R1[9], Rj[9];
        mat_mul(P.r, c_sc.jR[j], R1);
        V3 tp = xf_apply(P, c_sc.jt[j][0], c_sc.jt[j][1], c_sc.jt[j][2]);
        axis_angle(c_sc.jaxis[j], cq[j], sq[j], Rj);
        Xf& F = frames[1 + j];
        mat_mul(R1, Rj, F.r);
        F.t[0] = tp.x; F.t[1] = tp.y; F.t[2] = tp.z;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// 4. warp-cooperative GJK distance between two convex vertex sets held in shared memory
//    (restates what p.getClosestPoints computes on the margin-less cores; call sites ctlp.py:3267, :3300, :3353)
// ------------------------------------------------------------------------------------------------------------------
struct GjkCounters {
    unsigned calls, iters, dots;
};

__device__ __forceinline__ unsigned fkey(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// support vertex index of a hull in local direction d; the 32 lanes split the vertices
__device__ __forceinline__ int warp_support(const float4* __restrict__ v, int n, V3 d, int lane) {
    float best = -FLT_MAX;
    int bi = 0;
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

// closest point to the origin on segment ab; mask bit 0 = a kept, bit 1 = b kept
__device__ __forceinline__ V3 closest_segment(V3 a, V3 b, int& mask) {
    V3 ab = b - a;
    float t = -dot(a, ab), den = dot(ab, ab);
    if (t <= 0.0f || !(den > 0.0f)) { mask = 1; return a; }
    if (t >= den) { mask = 2; return b; }
    mask = 3;
    return a + (t / den) * ab;
}

// closest point to the origin on triangle abc (Ericson, Real-Time Collision Detection 5.1.5); mask = kept vertices.
// Degenerate (collinear) triangles fall back to the closest of the three edges instead of dividing by ~0.
__device__ __forceinline__ V3 closest_triangle(V3 a, V3 b, V3 c, int& mask) {
    V3 ab = b - a, ac = c - a;
    float d1 = -dot(ab, a), d2 = -dot(ac, a);
    if (d1 <= 0.0f && d2 <= 0.0f) { mask = 1; return a; }
    float d3 = -dot(ab, b), d4 = -dot(ac, b);
    if (d3 >= 0.0f && d4 <= d3) { mask = 2; return b; }
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f && d1 - d3 > 0.0f) { mask = 3; return a + (d1 / (d1 - d3)) * ab; }
    float d5 = -dot(ab, c), d6 = -dot(ac, c);
    if (d6 >= 0.0f && d5 <= d6) { mask = 4; return c; }
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f && d2 - d6 > 0.0f) { mask = 5; return a + (d2 / (d2 - d6)) * ac; }
    float va = d3 * d6 - d5 * d4;
    float e1 = d4 - d3, e2 = d5 - d6;
    if (va <= 0.0f && e1 >= 0.0f && e2 >= 0.0f && e1 + e2 > 0.0f) { mask = 6; return b + (e1 / (e1 + e2)) * (c - b); }
    float sum = va + vb + vc;
    V3 nrm = cross(ab, ac);
    // interior only if the triangle has a usable area: |n|^2 against the squared edge lengths
    if (sum > 0.0f && dot(nrm, nrm) > 1e-10f * dot(ab, ab) * dot(ac, ac)) {
        float denom = 1.0f / sum;
        mask = 7;
        return a + (vb * denom) * ab + (vc * denom) * ac;
    }
    int m1, m2, m3;
    V3 p1 = closest_segment(a, b, m1), p2 = closest_segment(a, c, m2), p3 = closest_segment(b, c, m3);
    float q1 = dot(p1, p1), q2 = dot(p2, p2), q3 = dot(p3, p3);
    if (q1 <= q2 && q1 <= q3) { mask = m1; return p1; }                                   // a=1, b=2
    if (q2 <= q3) { mask = (m2 & 1) | ((m2 & 2) << 1); return p2; }                       // a=1, c=4
    mask = m3 << 1;                                                                       // b=2, c=4
    return p3;
}

struct Simplex {
    V3 p0, p1, p2, p3;
    int i0, i1, i2, i3;  // (vertex of A << 16 | vertex of B) of each simplex point
    int n;
};

__device__ __forceinline__ void simplex_keep3(Simplex& S, V3 a, V3 b, V3 c, int ia, int ib, int ic, int mask) {
    int k = 0;
    if (mask & 1) { S.p0 = a; S.i0 = ia; k = 1; }
    if (mask & 2) { if (k == 0) { S.p0 = b; S.i0 = ib; } else { S.p1 = b; S.i1 = ib; } ++k; }
    if (mask & 4) {
        if (k == 0) { S.p0 = c; S.i0 = ic; } else if (k == 1) { S.p1 = c; S.i1 = ic; } else { S.p2 = c; S.i2 = ic; }
        ++k;
    }
    S.n = k;
}

// closest point of the simplex to the origin; reduces the simplex to the supporting face.  true = origin enclosed
__device__ __forceinline__ bool simplex_solve(Simplex& S, V3& v) {
    if (S.n == 1) { v = S.p0; return false; }
    if (S.n == 2) {
        V3 ab = S.p1 - S.p0;
        float t = -dot(S.p0, ab), den = dot(ab, ab);
        if (t <= 0.0f || den <= 0.0f) { v = S.p0; S.n = 1; return false; }
        if (t >= den) { v = S.p1; S.p0 = S.p1; S.i0 = S.i1; S.n = 1; return false; }
        v = S.p0 + (t / den) * ab;
        return false;
    }
    if (S.n == 3) {
        int mask;
        v = closest_triangle(S.p0, S.p1, S.p2, mask);
        simplex_keep3(S, S.p0, S.p1, S.p2, S.i0, S.i1, S.i2, mask);
        return false;
    }
    // tetrahedron: faces (012|3) (013|2) (023|1) (123|0).  The closest boundary point is taken over all four faces;
    // the origin counts as enclosed only if every face test says "inside" AND the tetrahedron is not flat -- in
    // float32 a sliver of four nearly coplanar support points must never certify a penetration.
    V3 A = S.p0, B = S.p1, Cc = S.p2, D = S.p3;
    int ia = S.i0, ib = S.i1, ic = S.i2, id = S.i3;
    float best = FLT_MAX;
    V3 bv = mk(0.f, 0.f, 0.f);
    int bmask = 0, bf = 0;
    bool inside_all = true;
#define SM_FACE(F, a, b, c, d)                                                  \
    {                                                                           \
        V3 nrm = cross(b - a, c - a);                                           \
        float sd = dot(d - a, nrm), so = -dot(a, nrm);                          \
        if (!(so * sd > 0.0f)) inside_all = false;                              \
        int m;                                                                  \
        V3 p = closest_triangle(a, b, c, m);                                    \
        float dd = dot(p, p);                                                   \
        if (dd < best) { best = dd; bv = p; bmask = m; bf = F; }                \
    }
    SM_FACE(0, A, B, Cc, D)
    SM_FACE(1, A, B, D, Cc)
    SM_FACE(2, A, Cc, D, B)
    SM_FACE(3, B, Cc, D, A)
#undef SM_FACE
    if (inside_all) {
        V3 e1 = B - A, e2 = Cc - A, e3 = D - A;
        float det = dot(e3, cross(e1, e2));
        float scale2 = dot(e1, e1) * dot(e2, e2) * dot(e3, e3);
        if (det * det > 1e-8f * scale2) return true;   // normalised volume above 1e-4: a genuine enclosure
    }
    if (bf == 0) simplex_keep3(S, A, B, Cc, ia, ib, ic, bmask);
    else if (bf == 1) simplex_keep3(S, A, B, D, ia, ib, id, bmask);
    else if (bf == 2) simplex_keep3(S, A, Cc, D, ia, ic, id, bmask);
    else simplex_keep3(S, B, Cc, D, ib, ic, id, bmask);
    v = bv;
    return false;
}

// Core distance between hull A (vertices vA in frame TA) and hull B.
//   upper   > 0: stop as soon as the distance is proven >= upper (returns a value >= upper): exact pruning of pairs
//                that cannot lower the running minimum / cannot be inside the query distance.
//   touch  >= 0: stop as soon as the distance is proven <= touch (returns a value <= touch): contact tests.
__device__ float gjk_warp(const float4* __restrict__ vA, int nA, const Xf& TA, const float4* __restrict__ vB, int nB,
                          const Xf& TB, V3 dir0, float upper, float touch, int lane, GjkCounters* cnt) {
    Simplex S;
    S.n = 0;
    S.i0 = S.i1 = S.i2 = S.i3 = -1;
    S.p0 = S.p1 = S.p2 = S.p3 = mk(0.f, 0.f, 0.f);
    V3 v = dir0;  // first search direction: from A towards B, so w = sA(-v') ... uses d = -v below with v = cA - cB
    float vv = dot(v, v);
    if (vv < 1e-12f) { v = mk(1.f, 0.f, 0.f); vv = 1.f; }
    bool have_point = false;
    if (cnt) cnt->calls++;
    for (int it = 0; it < 32; ++it) {
        V3 dA = xf_rot_t(TA, mk(-v.x, -v.y, -v.z));
        V3 dB = xf_rot_t(TB, v);
        int ia = warp_support(vA, nA, dA, lane);
        int ib = warp_support(vB, nB, dB, lane);
        if (cnt) { cnt->iters++; cnt->dots += (unsigned)(nA + nB); }
        float4 pa = vA[ia], pb = vB[ib];
        V3 w = xf_apply(TA, pa.x, pa.y, pa.z) - xf_apply(TB, pb.x, pb.y, pb.z);
        int id = (ia << 16) | ib;
        if (!have_point) {  // the first iteration only seeds the simplex with a real point of A - B
            S.p0 = w; S.i0 = id; S.n = 1;
            v = w; vv = dot(v, v);
            have_point = true;
            if (touch >= 0.0f && vv <= touch * touch) break;
            if (vv <= 1e-20f) return 0.0f;
            continue;
        }
        float vw = dot(v, w);
        if (upper > 0.0f && vw > 0.0f && vw * vw >= upper * upper * vv) return fmaxf(sqrtf(vv), upper);
        float nv = sqrtf(vv);
        if (vv - vw <= fmaxf(1e-6f * vv, 3e-7f * nv)) break;                     // converged
        if (id == S.i0 || id == S.i1 || id == S.i2 || id == S.i3) break;         // support already in the simplex
        if (S.n == 1) { S.p1 = w; S.i1 = id; }
        else if (S.n == 2) { S.p2 = w; S.i2 = id; }
        else { S.p3 = w; S.i3 = id; }
        S.n++;
        V3 nvv;
        if (simplex_solve(S, nvv)) return 0.0f;
        if (S.n < 4) S.i3 = -1;
        if (S.n < 3) S.i2 = -1;
        if (S.n < 2) S.i1 = -1;
        float nd = dot(nvv, nvv);
        if (!(nd < vv)) break;  // no progress (numerical floor) or a NaN from a degenerate simplex
        v = nvv; vv = nd;
        if (vv <= 1e-20f) return 0.0f;
        if (touch >= 0.0f && vv <= touch * touch) break;
    }
    return sqrtf(vv);
}
