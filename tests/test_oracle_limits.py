"""The oracle's safe-acceleration range (klimits restatement): closed forms, exactness and invariants."""
import numpy as np

from oracle import oracle

TS = 0.1


def rollout_check(scene, mode, episodes, steps, seed):
    rng = np.random.default_rng(seed)
    lo_p, hi_p = np.array(scene.pos_lo), np.array(scene.pos_hi)
    V, A, J = np.array(scene.vel_max), np.array(scene.acc_max), np.array(scene.jerk_max)
    worst = dict(p=0.0, v=0.0, a=0.0, j=0.0, codes=0)
    for _ in range(episodes):
        q = rng.uniform(lo_p * 0.9, hi_p * 0.9)
        v, a, u = np.zeros(7), np.zeros(7), np.ones(7)
        for st in range(steps):
            lo, hi, code = oracle.safe_range(scene, q[None], v[None], a[None])
            lo, hi, code = lo[0], hi[0], code[0]
            worst["codes"] += int((code != 0).sum())
            assert (hi >= lo).all()
            if mode == "rand":
                u = rng.uniform(-1, 1, 7)
            elif mode == "bang":
                u = np.sign(rng.uniform(-1, 1, 7))
            elif mode == "hold" and st % 10 == 0:
                u = np.sign(rng.uniform(-1, 1, 7))
            a1 = lo + 0.5 * (u + 1) * (hi - lo)
            tt = np.linspace(0, TS, 25)[:, None]
            jerk = (a1 - a) / TS
            aa, vv = a + jerk * tt, v + a * tt + 0.5 * jerk * tt ** 2
            qq = q + v * tt + 0.5 * a * tt ** 2 + jerk * tt ** 3 / 6
            worst["p"] = max(worst["p"], np.maximum(qq - hi_p, lo_p - qq).max())
            worst["v"] = max(worst["v"], (np.abs(vv) / V).max())
            worst["a"] = max(worst["a"], (np.abs(aa) / A).max())
            worst["j"] = max(worst["j"], (np.abs(jerk) / J).max())
            q, v, a = qq[-1], vv[-1], a1
    return worst


def test_limits_never_violated_under_random_and_adversarial_actions(space_scene):
    for mode, eps in (("rand", 40), ("max", 40), ("bang", 40), ("hold", 40)):
        w = rollout_check(space_scene, mode, eps, 60, 3)
        # the reference tolerates 1.001 (observations.py:379-403) and 1.002 for jerk (rewards.py:195-197)
        assert w["p"] < 1e-6 and w["v"] < 1 + 1e-9 and w["a"] < 1 + 1e-12 and w["j"] < 1 + 1e-9, (mode, w)
        if mode == "rand":
            assert w["codes"] == 0


def test_rest_state_has_full_range(space_scene):
    lo, hi, code = oracle.safe_range(space_scene, np.zeros((1, 7)), np.zeros((1, 7)), np.zeros((1, 7)))
    assert np.allclose(hi[0], space_scene.acc_max) and np.allclose(lo[0], -np.array(space_scene.acc_max))
    assert (code == 0).all()


def test_range_is_mirror_symmetric(space_scene):
    rng = np.random.default_rng(5)
    q = rng.uniform(space_scene.pos_lo, space_scene.pos_hi, (500, 7))
    v = rng.uniform(-1, 1, (500, 7)) * space_scene.vel_max
    a = rng.uniform(-1, 1, (500, 7)) * space_scene.acc_max
    lo, hi, code = oracle.safe_range(space_scene, q, v, a)
    lo2, hi2, code2 = oracle.safe_range(space_scene, -q, -v, -a)
    assert np.array_equal(lo, -hi2) and np.array_equal(hi, -lo2)
    assert np.array_equal(code == 0, code2 == 0)


def hardest_braking_peak_velocity(v, a, a1, J, A, n=4000):
    """Brute force: velocity peak when a goes linearly to a1 over TS and then ramps down by J*TS per knot (>= -A)."""
    peak, acc, vel = v, a, v
    nxt = a1
    for _ in range(8):
        t = np.linspace(0, TS, n)
        j = (nxt - acc) / TS
        vv = vel + acc * t + 0.5 * j * t * t
        peak = max(peak, vv.max())
        vel, acc = vv[-1], nxt
        if acc <= 0:
            break
        nxt = max(acc - J * TS, -A)
    return peak


def test_velocity_bound_is_exact(space_scene):
    """hi is feasible (hardest braking keeps v <= vmax) and hi + eps is not, whenever the velocity bound binds."""
    rng = np.random.default_rng(7)
    V, A, J = np.array(space_scene.vel_max), np.array(space_scene.acc_max), np.array(space_scene.jerk_max)
    checked = 0
    for _ in range(400):
        j = rng.integers(0, 7)
        v = rng.uniform(0.5, 0.999) * V[j]
        a = rng.uniform(-0.5, 1.0) * A[j]
        q = np.zeros(7)
        vv, aa = np.zeros(7), np.zeros(7)
        vv[j], aa[j] = v, a
        lo, hi, code = oracle.safe_range(space_scene, q[None], vv[None], aa[None])
        h, l = hi[0, j], lo[0, j]
        if code[0, j] != 0 or h >= min(A[j], a + J[j] * TS) - 1e-9 or h <= l + 1e-9:
            continue  # velocity bound not the binding one
        checked += 1
        assert hardest_braking_peak_velocity(v, a, h, J[j], A[j]) <= V[j] + 1e-7
        assert hardest_braking_peak_velocity(v, a, h + 1e-3 * A[j], J[j], A[j]) > V[j]
    assert checked > 50


def test_position_bound_is_exact(space_scene):
    rng = np.random.default_rng(8)
    P, V, A, J = np.array(space_scene.pos_hi), np.array(space_scene.vel_max), np.array(space_scene.acc_max), \
        np.array(space_scene.jerk_max)

    def peak_position(p, v, a, a1, j):
        best, acc, vel, pos, nxt = p, a, v, p, a1
        for _ in range(40):
            t = np.linspace(0, TS, 2000)
            jj = (nxt - acc) / TS
            pp = pos + vel * t + 0.5 * acc * t * t + jj * t ** 3 / 6
            best = max(best, pp.max())
            vel, pos, acc = vel + acc * TS + 0.5 * jj * TS * TS, pp[-1], nxt
            if vel <= 0 and acc <= 0:
                break
            nxt = max(acc - J[j] * TS, -A[j])
        return best
    checked = 0
    for _ in range(400):
        j = rng.integers(0, 7)
        p = P[j] - rng.uniform(0.0, 0.15)
        v = rng.uniform(0.0, 0.6) * V[j]
        a = rng.uniform(-0.5, 0.5) * A[j]
        q, vv, aa = np.zeros(7), np.zeros(7), np.zeros(7)
        q[j], vv[j], aa[j] = p, v, a
        lo, hi, code = oracle.safe_range(space_scene, q[None], vv[None], aa[None])
        if code[0, j] != 0 or not (code[0, j] == 0 and hi[0, j] < min(A[j], a + J[j] * TS) - 1e-6):
            continue
        # was it the position bound?  compare with the range of the same state far from the limit
        q2 = q.copy()
        q2[j] = 0.0
        _, hi_free, _ = oracle.safe_range(space_scene, q2[None], vv[None], aa[None])
        if hi_free[0, j] - hi[0, j] < 1e-6 or hi[0, j] <= lo[0, j] + 1e-9:
            continue
        checked += 1
        assert peak_position(p, v, a, hi[0, j], j) <= P[j] + 1e-7
        assert peak_position(p, v, a, hi[0, j] + 1e-3 * A[j], j) > P[j]
    assert checked > 30


def test_invalid_states_get_violation_codes(space_scene):
    # at the velocity limit with maximum acceleration no admissible jerk can stop the overshoot
    v, a = np.zeros((1, 7)), np.zeros((1, 7))
    v[0, 0], a[0, 0] = space_scene.vel_max[0], space_scene.acc_max[0]
    _, _, code = oracle.safe_range(space_scene, np.zeros((1, 7)), v, a)
    assert code[0, 0] != 0 and (code[0, 1:] == 0).all()
    # at the position limit while still moving outwards
    q = np.zeros((1, 7))
    q[0, 2] = space_scene.pos_hi[2]
    v = np.zeros((1, 7))
    v[0, 2] = 1.0
    _, _, code = oracle.safe_range(space_scene, q, v, np.zeros((1, 7)))
    assert code[0, 2] & 4


def test_interpolation_matches_reference_twins(space_scene):
    """One oracle step with a frozen scene: the new knot equals the formulas of actions.py:468-487 at t = dt."""
    from safemotionsrisk_b200 import space_backup_config
    from safemotionsrisk_b200.scene import Scene
    scene = Scene(space_backup_config(contact_check_stride=0))
    rng = np.random.default_rng(2)
    n = 16
    env = oracle.OracleEnvs(scene, n)
    q = rng.uniform(-0.5, 0.5, (n, 7))
    v = rng.uniform(-0.3, 0.3, (n, 7))
    a = rng.uniform(-1, 1, (n, 7))
    ob = np.zeros((n, 16))
    env.set_state(q, v, a, ob)
    lo, hi, _ = oracle.safe_range(scene, q, v, a)
    u = rng.uniform(-1, 1, (n, 7)).astype(np.float32)
    env.step(u)
    a1 = lo + 0.5 * (u.astype(np.float64) + 1) * (hi - lo)
    t = TS
    q1 = q + v * t + 0.5 * a * t ** 2 + 1 / 6 * ((a1 - a) / TS) * t ** 3
    v1 = v + a * t + 0.5 * ((a1 - a) / TS) * t ** 2
    assert np.allclose(env.kin[:, 16:23], a1, rtol=0, atol=1e-12)
    assert np.allclose(env.kin[:, 0:7], q1, rtol=0, atol=1e-12)
    assert np.allclose(env.kin[:, 8:15], v1, rtol=0, atol=1e-12)
    # v1 = v0 + (a0 + a1) dt / 2 and p1 = p0 + v0 dt + (a0/3 + a1/6) dt^2 (klimits model)
    assert np.allclose(env.kin[:, 8:15], v + 0.5 * (a + a1) * TS, atol=1e-12)
    assert np.allclose(env.kin[:, 0:7], q + v * TS + (a / 3 + a1 / 6) * TS ** 2, atol=1e-12)


def test_peak_derivative_matches_finite_differences():
    """The position-bound solve is a Newton iteration on peak(a1) - limit with the derivative carried along the
    simulated braking profile (oracle/smenv_oracle.c: pos_peak_d): away from the kinks of the profile the derivative
    agrees with a central difference, and the peak is non-decreasing in a1."""
    import ctypes as C
    lib = oracle.lib()
    lib.smo_pos_peak.restype = C.c_double
    lib.smo_pos_peak.argtypes = [C.c_double] * 7 + [C.POINTER(C.c_double)]
    rng = np.random.default_rng(5)
    ts, checked = 0.1, 0
    for _ in range(4000):
        J, A = rng.choice([150.0, 200.0, 300.0, 400.0]), rng.choice([7.5, 10.0, 15.0, 20.0])
        p, v, a = rng.uniform(-2.5, 2.5), rng.uniform(-1.7, 1.7), rng.uniform(-A, A)
        a1 = rng.uniform(max(-A, a - J * ts), min(A, a + J * ts))
        d = C.c_double()
        f0 = lib.smo_pos_peak(p, v, a, a1, J, A, ts, C.byref(d))
        h = 1e-5
        dm, dp_ = C.c_double(), C.c_double()
        fm = lib.smo_pos_peak(p, v, a, a1 - h, J, A, ts, C.byref(dm))
        fp = lib.smo_pos_peak(p, v, a, a1 + h, J, A, ts, C.byref(dp_))
        assert fp >= fm - 1e-12                       # monotone
        assert f0 >= p                                # the peak is at least the start position
        if abs(dm.value - dp_.value) > 1e-6 * max(1.0, abs(d.value)):
            continue                                  # a kink of the profile between a1 - h and a1 + h
        fd = (fp - fm) / (2 * h)
        assert abs(fd - d.value) <= 1e-6 + 1e-4 * abs(d.value), (p, v, a, a1, fd, d.value)
        checked += 1
    assert checked > 2000
