// smenv_joint.cuh -- joint-space half of the env step: one THREAD per (env, joint), float64, strict IEEE.
//
//   safe acceleration range  -> action mapping -> 24 interpolated setpoints + motor-tracked pose -> new knot
//   (actions.py:206-211, :268-280, :412-443; safe_motions_base.py:1179-1185, :1233-1277)
//
// Eight lanes serve one env (lane 7 idles for the 7-joint iiwa), so a warp advances four envs and ~88 % of the FP64
// lanes do useful work; the geometry kernel (smenv_step.cuh, one warp per env) then consumes the per-env scratch
// record written here.  Splitting the step keeps each kernel's code inside the instruction cache: the fused version
// stalled ~48 cycles per issue on instruction fetch (profiles/r01_fused_step_ball.txt).
#pragma once
#include "smenv_kernels.cuh"

#define SM_SCRATCH_FLOATS (2 * SM_MAX_SUB * SM_MAX_JOINTS + 16) /* qsub[32][8] + misc[16] + qset[32][8] */
#define SM_QSET_OFF (SM_MAX_SUB * SM_MAX_JOINTS + 16) /* setpoint pose of every sub-step (target points only) */
#define SM_MISC_OFF (SM_MAX_SUB * SM_MAX_JOINTS)
#define SM_MISC_RCODE 0     /* joint_kernel: OR of the violation codes */
#define SM_MISC_JERK 1      /* joint_kernel: max relative jerk */
#define SM_MISC_UMAX 2      /* joint_kernel: max |action| */
#define SM_MISC_MASK0 4     /* contact_broad_kernel: bit k = sub-step k+1 needs a narrow-phase test, obstacle 0 */
#define SM_MISC_MASK1 5     /* idem obstacle 1 */
#define SM_MISC_HIT 6       /* contact_narrow_kernel: 1-based sub-step of the first contact, 0 = none */
#define SM_MISC_DSTATIC 8   /* distance_kernel */
#define SM_MISC_DSELF 9
#define SM_MISC_DMOVING 10

struct JointArgs {
    SmBuffers buf;
    int n;
    int random_actions;
    uint32_t k0, k1, step_counter;
    int env_base;    // index of env 0 of this launch in the whole vector env (Philox counters use the global index)
    float* scratch;  // [n][SM_SCRATCH_FLOATS]
    int* worklist;   // [0] = GJK item counter, [1] = overflowed items of the step: both cleared here
    unsigned long long* counters;  // device SmCounters in the counting mode, else NULL
    int* cwork;      // [0] = number of envs listed for the fine contact planning: cleared here
    int* tasks;      // [0] = number of position bounds to solve (cleared here), [1..] = heavy index * 2 + side
    double* hpar;    // [8 n][SM_HPAR] hand-over records of the deferred instances
    int* heavy;      // [0] = number of (env, joint) instances whose position bound needs the iterative solve,
                     // [1..] = env * 8 + joint (filled by joint_kernel, consumed by joint_heavy_kernel)
    int set;            // 0: the robot's joints; 1: the human's (limits c_sc.hu.lim, kin = buf.hkin, actions = buf.hactions)
    int nj;             // joints of the set
    int defer;          // human: do not advance, write (lo, hi, mapped acceleration, code) of every joint to range_out; the
                        // braking-trajectory check decides the end acceleration first (actions.py:343-376)
    double* range_out;  // [n][8][4]
    double track_vel;   // 0.87 with use_controller_target_velocities, else 0 (robot_scene_base.py:792-805)
    int store_qset;     // also store the setpoint pose of every sub-step (target points)
    int* clear_extra;   // a further list counter cleared by joint_kernel (human: the pose units of the braking check)
    int keep_overflow;  // do not clear the overflow count of the item list (the human's braking-trajectory items came first)
    const float* exec;  // [n][n_joints] actions to execute when the risk gate is on (the backup policy's where the
                        // proposed action was rated risky), else NULL: buf.actions are executed.  The action punishment
                        // always rates the proposed action (safe_motions_base.py:1066, actions.py:328-331)
};

// the action the motors execute: the gated one if the risk gate is on
__device__ __forceinline__ float joint_exec_action(const JointArgs& A, int env, int j, float proposed) {
    return A.exec ? A.exec[(size_t)env * A.nj + j] : proposed;
}
__device__ __forceinline__ const JointLim& joint_limits(const JointArgs& A) { return A.set ? c_sc.hu.lim : c_sc.lim; }
__device__ __forceinline__ double* joint_kin(const JointArgs& A, size_t env) {
    return (A.set ? A.buf.hkin : A.buf.kin) + env * SM_KIN_STRIDE;
}
__device__ __forceinline__ float joint_action(const JointArgs& A, int env, int j) {
    if (A.random_actions) {  // get_random_action (safe_motions_base.py:1327-1328)
        uint4 r = philox((uint32_t)(env + A.env_base), A.step_counter, (uint32_t)j, 0xAC71u, A.k0, A.k1);
        return 2.0f * u01f(r.x) - 1.0f;
    }
    return (A.set ? A.buf.hactions : A.buf.actions)[(size_t)env * A.nj + j];
}

// Action mapping, the S interpolated setpoints with the motor-tracked pose, the new knot (actions.py:268-280,
// :412-443; safe_motions_base.py:1179-1185, :1233-1277).  Returns the relative jerk of the step (rewards.py:181-186).
// setpoint of sub-step k (1-based) of a step from (q, v, a) to the end acceleration a1: the arithmetic of the loop in
// joint_advance_a1, operation by operation (the human's target-point check recomputes setpoints instead of storing them)
__device__ __forceinline__ double joint_setpoint_at(double q, double v, double a, double a1, int k) {
    const double jerk = xdiv(xsub(a1, a), c_sc.ts);
    const double ha = xmul(0.5, a), sj = xmul(1.0 / 6.0, jerk);
    const double tk = c_sc.sub_t[k];
    return xadd(xadd(xadd(q, xmul(v, tk)), xmul(xmul(ha, tk), tk)), xmul(xmul(xmul(sj, tk), tk), tk));
}
__device__ __forceinline__ float joint_advance_a1(double* kin, float* scr, int j, double q, double v, double a, double qa,
                                                  double a1, double track_vel, bool store_qset, double jerk_max,
                                                  float* dq_out = nullptr /* max_k |setpoint(k) - setpoint(S)| */) {
    const int S = c_sc.substeps;
    const double dt = xdiv(c_sc.ts, (double)S);
    const double tvdt = xmul(track_vel, dt);
    const double jerk = xdiv(xsub(a1, a), c_sc.ts);      // actions.py:468-487, hoisted out of the sub-step loop
    const double hj = xmul(0.5, jerk), ha = xmul(0.5, a), sj = xmul(1.0 / 6.0, jerk);
    double q1 = q, v1 = v;
    float qmin = FLT_MAX, qmax = -FLT_MAX;
    for (int k = 1; k <= S; ++k) {
        const double tk = c_sc.sub_t[k];
        double vs = xadd(xadd(v, xmul(a, tk)), xmul(xmul(hj, tk), tk));
        double qs = xadd(xadd(xadd(q, xmul(v, tk)), xmul(xmul(ha, tk), tk)), xmul(xmul(xmul(sj, tk), tk), tk));
        scr[(k - 1) * SM_MAX_JOINTS + j] = (float)qa;    // pose seen by the collision detection of sub-step k
        if (store_qset) scr[SM_QSET_OFF + (k - 1) * SM_MAX_JOINTS + j] = (float)qs;  // ctlp.py:2787-2791
        qa = xadd(xadd(qa, xmul(c_sc.track_kp, xsub(qs, qa))), xmul(tvdt, vs));
        q1 = qs; v1 = vs;                                // k == S: the new knot
        qmin = fminf(qmin, (float)qs); qmax = fmaxf(qmax, (float)qs);
    }
    kin[j] = q1; kin[8 + j] = v1; kin[16 + j] = a1; kin[24 + j] = qa;
    if (dq_out) *dq_out = fmaxf(qmax - (float)q1, (float)q1 - qmin);
    return (float)(fabs(jerk) / jerk_max);
}
__device__ __forceinline__ float joint_advance(const JointArgs& A, double* kin, float* scr, int j, double q, double v,
                                               double a, double qa, double lo, double hi, float uf) {
    const double a1 = map_action((double)uf, lo, hi);
    if (A.defer) {   // human: the braking-trajectory check comes first
        double* ro = A.range_out + ((size_t)(kin - (A.set ? A.buf.hkin : A.buf.kin)) / SM_KIN_STRIDE * 8 + j) * 4;
        ro[0] = lo; ro[1] = hi; ro[2] = a1;
        return 0.0f;
    }
    return joint_advance_a1(kin, scr, j, q, v, a, qa, a1, A.track_vel, A.store_qset != 0, joint_limits(A).jerk_max[j]);
}

// First pass: every (env, joint) whose position bounds are certainly inactive (the common case) is finished here;
// the others are compacted into the `heavy` list so that the iterative solve runs with full warps afterwards
// instead of stalling 31 idle lanes (the single-pass kernel averaged 5.6 active lanes per instruction).
#ifndef JK_MIN_BLOCKS
#define JK_MIN_BLOCKS 1   /* resident CTAs per SM the register allocation aims for (4: 64 registers, no change: FP64-pipe bound) */
#endif
__global__ void __launch_bounds__(256, JK_MIN_BLOCKS) joint_kernel(JointArgs A) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int env = t >> 3, j = t & 7;
    const int nj = A.nj;
    const bool valid = env < A.n;
    const bool jl = valid && j < nj;
    const size_t e = valid ? (size_t)env : 0;
    double* kin = joint_kin(A, e);
    const JointLim& L = joint_limits(A);
    float* scr = A.scratch + e * SM_SCRATCH_FLOATS;
    float uf = 0.0f, jerk_rel = 0.0f;
    int code = 0;
    bool defer = false;
    if (jl) {
        const double q = kin[j], v = kin[8 + j], a = kin[16 + j], qa = kin[24 + j];
        uf = joint_action(A, env, j);
        double lo, hi;
        safe_range_light(L, j, q, v, a, lo, hi, code, defer);
        if (!defer) jerk_rel = joint_advance(A, kin, scr, j, q, v, a, qa, lo, hi, joint_exec_action(A, env, j, uf));
        else code = 0;  // reported by joint_heavy_kernel
    } else if (valid) {
        for (int k = 0; k < c_sc.substeps; ++k) scr[k * SM_MAX_JOINTS + j] = 0.0f;
    }
    // warp-aggregated append of the deferred instances
    const unsigned dm = __ballot_sync(FULL, defer);
    if (dm) {
        int base = 0;
        if (lane == __ffs(dm) - 1) base = atomicAdd(A.heavy, __popc(dm));
        base = __shfl_sync(FULL, base, __ffs(dm) - 1);
        if (defer) A.heavy[1 + base + __popc(dm & ((1u << lane) - 1u))] = t;
    }
    // per-env reductions over the eight lanes of the env (xor 1, 2, 4 stay inside the group)
    unsigned rc = (unsigned)code;
    float um = fabsf(uf);
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) {
        rc |= __shfl_xor_sync(FULL, rc, m);
        jerk_rel = fmaxf(jerk_rel, __shfl_xor_sync(FULL, jerk_rel, m));
        um = fmaxf(um, __shfl_xor_sync(FULL, um, m));
    }
    if (valid) {  // misc[0..7]: codes (integer bits) / jerk / max action; contact masks and hit cleared
        float val = j == SM_MISC_RCODE ? __int_as_float((int)rc) : j == SM_MISC_JERK ? jerk_rel : j == SM_MISC_UMAX ? um : 0.0f;
        scr[SM_MISC_OFF + j] = val;
    }
    if (t == 0 && A.worklist) { A.worklist[0] = 0; if (!A.keep_overflow) A.worklist[1] = 0; }
    if (t == 0 && A.cwork) A.cwork[0] = 0;
    if (t == 0 && A.tasks) A.tasks[0] = 0;
    if (t == 0 && A.clear_extra) A.clear_extra[0] = 0;
}

// Second pass over the deferred (env, joint) instances, as three small kernels so that each runs with densely
// packed lanes (a single kernel averaged 5.3 active lanes per instruction: a third of the joints are deferred, a
// tenth need the iterative solve, and those need it for very different numbers of iterations):
//   joint_first_kernel   thread = deferred instance: the cheap bounds again and the FIRST evaluation of both position
//                        bounds (is the bound active at all?); active bounds are appended to the task list
//   joint_solve_kernel   thread = task: the iterative solve (pos_upper_rest)
//   joint_final_kernel   thread = deferred instance: clamp, map the action, advance
// Per-instance hand-over record (SM_HPAR doubles): lo, hi, code, fr_hi, fr_lo, res_hi, res_lo.
#define SM_HPAR 8
#define SM_HEAVY_THREADS 128

#ifndef JF_MIN_BLOCKS
#define JF_MIN_BLOCKS 1   /* resident CTAs per SM the register allocation aims for (8: 64 registers, no change) */
#endif
__global__ void __launch_bounds__(SM_HEAVY_THREADS, JF_MIN_BLOCKS) joint_first_kernel(JointArgs A) {
    const int n_heavy = A.heavy[0];
    const int lane = threadIdx.x & 31;
    const double ts = c_sc.ts;
#pragma unroll 1
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < n_heavy; base += gridDim.x * blockDim.x) {
        const int i = base + lane;
        bool t_hi = false, t_lo = false;
        if (i < n_heavy) {
            const int t = A.heavy[1 + i];
            const int env = t >> 3, j = t & 7;
            const double* kin = joint_kin(A, (size_t)env);
            const JointLim& L = joint_limits(A);
            const double q = kin[j], v = kin[8 + j], a = kin[16 + j];
            double lo, hi;
            int code;
            bool need_pos;
            safe_range_light(L, j, q, v, a, lo, hi, code, need_pos);
            const double J = L.jerk_max[j], Am = L.acc_max[j];
            const double fr_hi = pos_upper_first(q, v, a, L.pos_hi[j], hi, J, Am, ts);
            const double fr_lo = pos_upper_first(-q, -v, -a, -L.pos_lo[j], -lo, J, Am, ts);
            double* hp = A.hpar + (size_t)i * SM_HPAR;
            hp[0] = lo; hp[1] = hi; hp[2] = (double)code; hp[3] = fr_hi; hp[4] = fr_lo; hp[5] = SM_BIG; hp[6] = SM_BIG;
            t_hi = fr_hi > 0.0;
            t_lo = fr_lo > 0.0;
        }
        const unsigned mh = __ballot_sync(FULL, t_hi), ml = __ballot_sync(FULL, t_lo);
        const int total = __popc(mh) + __popc(ml);
        if (total) {
            int b0 = 0;
            if (lane == 0) b0 = atomicAdd(A.tasks, total);
            b0 = __shfl_sync(FULL, b0, 0);
            const unsigned below = (1u << lane) - 1u;
            if (t_hi) A.tasks[1 + b0 + __popc(mh & below)] = 2 * i;
            if (t_lo) A.tasks[1 + b0 + __popc(mh) + __popc(ml & below)] = 2 * i + 1;
        }
    }
    if (A.counters && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&A.counters[8], (unsigned long long)n_heavy);
}

// The solves are nested loops of very different lengths (up to 40 Newton steps, each simulating a braking profile of
// 1..16 intervals), and a warp pays for the union of its lanes' paths: run task by task, a warp averaged 6.7 active
// lanes.  Here the nest is flattened into a state machine whose single loop body is ONE interval of the braking
// profile (the body of pos_peak_impl<true>), executed by all lanes in lockstep; a lane whose profile ends steps its
// root search, a lane whose task ends takes the next task of the warp's chunk.  The operations of every task are
// exactly those of pos_upper_rest / pos_peak_d, in the same order (bit-identical results).
#ifndef SM_SOLVE_CHUNK_MIN
#define SM_SOLVE_CHUNK_MIN 32
#endif
#ifndef JS_MIN_BLOCKS
#define JS_MIN_BLOCKS 1   /* resident CTAs per SM the register allocation aims for (8: 64 registers and 120 B of spill, 1.5 % slower) */
#endif
__global__ void __launch_bounds__(SM_HEAVY_THREADS, JS_MIN_BLOCKS) joint_solve_kernel(JointArgs A) {
    const int n_task = A.tasks[0];
    const double ts = c_sc.ts;
    const int lane = threadIdx.x & 31;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    int chunk = (n_task + n_warps - 1) / n_warps;
    if (chunk < SM_SOLVE_CHUNK_MIN) chunk = SM_SOLVE_CHUNK_MIN;  // enough tasks per warp to keep its lanes refilled
    int cursor = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * chunk;
    const int end = cursor + chunk < n_task ? cursor + chunk : n_task;
    // per-lane task state
    bool busy = false;
    double* out = nullptr;
    double P0 = 0, V0 = 0, A0 = 0, pmax = 0, lo = 0, J = 1, Am = 1;   // the task (mirrored for a lower bound)
    double xl = 0, xr = 0, fr = 0, x = 0;                              // bracket, current iterate
    int iit = 0, phase = 0;                                            // phase 0: evaluating f(lo)
    double p = 0, v = 0, a = 0, an = 0, best = 0;                      // braking profile under evaluation
    double dp = 0, dv = 0, da = 0, dan = 1, dbest = 0;                 // its derivative with respect to the iterate
    int pit = 0;
    unsigned n_intervals = 0;   // braking-profile intervals this lane evaluated (counting mode: algorithmic flops of the solve)
#pragma unroll 1
    while (true) {
        const unsigned idle = __ballot_sync(FULL, !busy);
        if (cursor < end && (idle == FULL || __popc(idle) >= 8)) {
            const int my = cursor + __popc(idle & ((1u << lane) - 1u));
            cursor += __popc(idle);
            if (!busy && my < end) {
                const int task = A.tasks[1 + my], i = task >> 1, sd = task & 1;
                const int t = A.heavy[1 + i];
                const int env = t >> 3, j = t & 7;
                const double* kin = joint_kin(A, (size_t)env);
                const JointLim& L = joint_limits(A);
                double* hp = A.hpar + (size_t)i * SM_HPAR;
                const double sg = sd == 0 ? 1.0 : -1.0;
                P0 = sg * kin[j]; V0 = sg * kin[8 + j]; A0 = sg * kin[16 + j];
                pmax = sd == 0 ? L.pos_hi[j] : -L.pos_lo[j];
                lo = sd == 0 ? hp[0] : -hp[1];
                xr = sd == 0 ? hp[1] : -hp[0];        // hi
                fr = hp[3 + sd];
                J = L.jerk_max[j]; Am = L.acc_max[j];
                out = hp + 5 + sd;
                phase = 0; x = lo;
                p = P0; v = V0; a = A0; an = x; best = P0; pit = 0;
                dp = 0.0; dv = 0.0; da = 0.0; dan = 1.0; dbest = 0.0;
                busy = true;
            }
        }
        if (!__any_sync(FULL, busy)) break;
        if (!busy) continue;
        // ---------------- one interval of the braking profile (body of pos_peak_impl<true>)
        bool eval_done = false;
        ++n_intervals;
        {
            const double j = xdiv(xsub(an, a), ts);
            const double dj = xdiv(xsub(dan, da), ts);
            double tau = -1.0;
            if (j == 0.0) {
                if (a < 0.0 && v > 0.0) tau = xdiv(-v, a);
            } else {
                const double disc = xsub(xmul(a, a), xmul(xmul(2.0, j), v));
                if (disc >= 0.0) {
                    const double s = xsqrt(disc);
                    if (a <= 0.0) {
                        if (xsub(s, a) > 0.0) tau = xdiv(xmul(2.0, v), xsub(s, a));
                    } else {
                        tau = xdiv(xsub(-a, s), j);
                    }
                }
            }
            if (tau > 0.0 && tau <= ts) {
                const double pk = xadd(xadd(xadd(p, xmul(v, tau)), xmul(xmul(xmul(0.5, a), tau), tau)),
                                       xdiv(xmul(xmul(xmul(j, tau), tau), tau), 6.0));
                if (pk > best) {
                    best = pk;
                    dbest = xadd(xadd(xadd(dp, xmul(dv, tau)), xmul(xmul(xmul(0.5, da), tau), tau)),
                                 xdiv(xmul(xmul(xmul(dj, tau), tau), tau), 6.0));
                }
            }
            const double pn = xadd(xadd(p, xmul(v, ts)), xmul(xmul(xadd(xdiv(a, 3.0), xdiv(an, 6.0)), ts), ts));
            const double vn = xadd(v, xmul(xmul(xadd(a, an), ts), 0.5));
            const double dpn = xadd(xadd(dp, xmul(dv, ts)), xmul(xmul(xadd(xdiv(da, 3.0), xdiv(dan, 6.0)), ts), ts));
            const double dvn = xadd(dv, xmul(xmul(xadd(da, dan), ts), 0.5));
            p = pn; v = vn; a = an;
            dp = dpn; dv = dvn; da = dan;
            if (p > best) { best = p; dbest = dp; }
            if (a <= -Am) {
                if (v > 0.0) {
                    const double pk = xadd(p, xdiv(xmul(v, v), xmul(2.0, Am)));
                    if (pk > best) { best = pk; dbest = xadd(dp, xdiv(xmul(v, dv), Am)); }
                }
                eval_done = true;
            } else if (v <= 0.0 && a <= 0.0) {
                eval_done = true;
            } else {
                an = xsub(a, xmul(J, ts));
                if (an < -Am) { an = -Am; dan = 0.0; }
                if (++pit >= 16) eval_done = true;
            }
        }
        if (!eval_done) continue;
        // ---------------- the profile ended: one step of pos_upper_rest
        const double f = xsub(best, pmax);
        bool task_done = false;
        double result = 0.0;
        double xn = 0.0;
        if (phase == 0) {
            const double fl = f;
            if (fl > 0.0) { result = fl > 1e-6 ? -SM_BIG : lo; task_done = true; }
            else {
                xl = lo; iit = 0; phase = 1;
                xn = xsub(xr, xdiv(xmul(fr, xsub(xr, xl)), xsub(fr, fl)));   // first iterate: secant of the bracket
                if (!(xn > xl && xn < xr)) xn = xmul(0.5, xadd(xl, xr));
            }
        } else if (f <= 0.0 && f > -SM_POS_SOLVE_TOL) {
            result = x; task_done = true;   // safe and within 1e-10 rad of the limit
        } else {
            if (f <= 0.0) xl = x; else xr = x;
            if (xsub(xr, xl) <= 1e-9 || ++iit >= 40) { result = xl; task_done = true; }
            else {
                xn = xl;
                if (dbest > 0.0) xn = xsub(x, xdiv(xadd(f, 0.5 * SM_POS_SOLVE_TOL), dbest));   // Newton step
                if (!(xn > xl && xn < xr)) xn = xmul(0.5, xadd(xl, xr));
            }
        }
        if (task_done) { *out = result; busy = false; }
        else {
            x = xn;
            p = P0; v = V0; a = A0; an = x; best = P0; pit = 0;
            dp = 0.0; dv = 0.0; da = 0.0; dan = 1.0; dbest = 0.0;
        }
    }
    if (A.counters && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&A.counters[9], (unsigned long long)n_task);
    if (A.counters) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_intervals += __shfl_xor_sync(FULL, n_intervals, o);
        if (lane == 0 && n_intervals) atomicAdd(&A.counters[7], (unsigned long long)n_intervals);
    }
}

__global__ void __launch_bounds__(SM_HEAVY_THREADS) joint_final_kernel(JointArgs A) {
    const int n_heavy = A.heavy[0];
#pragma unroll 1
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_heavy; i += gridDim.x * blockDim.x) {
        const int t = A.heavy[1 + i];
        const int env = t >> 3, j = t & 7;
        double* kin = joint_kin(A, (size_t)env);
        float* scr = A.scratch + (size_t)env * SM_SCRATCH_FLOATS;
        const double q = kin[j], v = kin[8 + j], a = kin[16 + j], qa = kin[24 + j];
        const double* hp = A.hpar + (size_t)i * SM_HPAR;
        double lo = hp[0], hi = hp[1];
        int code = (int)hp[2];
        if (c_sc.limit_position) clamp_range(lo, hi, -hp[6], hp[5], CODE_POS_HI, CODE_POS_LO, code);
        const float uf = joint_exec_action(A, env, j, joint_action(A, env, j));
        const float jerk_rel = joint_advance(A, kin, scr, j, q, v, a, qa, lo, hi, uf);
        // non-negative floats order like their bit patterns
        atomicMax(reinterpret_cast<int*>(scr + SM_MISC_OFF + SM_MISC_JERK), __float_as_int(jerk_rel));
        if (code) atomicOr(reinterpret_cast<int*>(scr + SM_MISC_OFF + SM_MISC_RCODE), code);
    }
}
