"""Oracle FK and GJK against independent computations (closed forms, brute-force QP)."""
import numpy as np
from scipy.optimize import minimize

from oracle import oracle

IDENT = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float64)


def rand_xf(rng, scale=0.3):
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] *= -1
    return np.concatenate([q.reshape(-1), rng.uniform(-scale, scale, 3)])


def qp_distance(a, b):
    na, nb = len(a), len(b)
    m = np.concatenate([a, -b]).T
    cons = [{"type": "eq", "fun": lambda x: x[:na].sum() - 1,
             "jac": lambda x: np.concatenate([np.ones(na), np.zeros(nb)])},
            {"type": "eq", "fun": lambda x: x[na:].sum() - 1,
             "jac": lambda x: np.concatenate([np.zeros(na), np.ones(nb)])}]
    x0 = np.concatenate([np.ones(na) / na, np.ones(nb) / nb])
    r = minimize(lambda x: ((m @ x) ** 2).sum(), x0, jac=lambda x: 2 * m.T @ (m @ x), bounds=[(0, 1)] * (na + nb),
                 constraints=cons, method="SLSQP", options=dict(maxiter=500, ftol=1e-16))
    return np.sqrt(max(r.fun, 0.0))


def test_fk_zero_pose_matches_urdf_chain(space_scene):
    # robot.urdf:4-231: world -> (-0.2,0,0) adapter -> (0,0,0.02) link0 -> J1 z@0.157 -> J2 @0.183 -> J3 @0.185 ->
    # J4 @0.215 -> J5 @0.4 -> J6 @0 -> J7 @0 ; all rpy = 0
    fr = oracle.fk(space_scene, np.zeros(7))
    z = np.cumsum([0.02 + 0.157, 0.183, 0.185, 0.215, 0.4, 0.0, 0.0])
    for j in range(7):
        assert np.allclose(fr[1 + j][9:], [-0.2, 0, z[j]], atol=1e-12)
        assert np.allclose(fr[1 + j][:9], IDENT[:9], atol=1e-12)


def test_fk_single_joint_rotations(space_scene):
    q = np.zeros(7)
    q[0] = 0.7                                   # J1 about +z
    fr = oracle.fk(space_scene, q)
    c, s = np.cos(0.7), np.sin(0.7)
    assert np.allclose(fr[1][:9].reshape(3, 3), [[c, -s, 0], [s, c, 0], [0, 0, 1]])
    q = np.zeros(7)
    q[3] = 0.5                                   # J4 about -y
    fr = oracle.fk(space_scene, q)
    c, s = np.cos(-0.5), np.sin(-0.5)
    assert np.allclose(fr[4][:9].reshape(3, 3), [[c, 0, s], [0, 1, 0], [-s, 0, c]])
    # the elbow bends the forearm: link 5 origin = link4 origin + R_y(-0.5) (0,0,0.4)
    assert np.allclose(fr[5][9:], fr[4][9:] + fr[4][:9].reshape(3, 3) @ [0, 0, 0.4])


def test_fk_is_a_rigid_chain(space_scene):
    rng = np.random.default_rng(0)
    for _ in range(20):
        fr = oracle.fk(space_scene, rng.uniform(space_scene.pos_lo, space_scene.pos_hi))
        for f in fr:
            r = f[:9].reshape(3, 3)
            assert np.allclose(r @ r.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(r) - 1) < 1e-12
        for j, d in enumerate([0.183, 0.185, 0.215, 0.4, 0.0, 0.0]):   # link lengths are pose independent
            assert abs(np.linalg.norm(fr[2 + j][9:] - fr[1 + j][9:]) - d) < 1e-12


def test_gjk_analytic_boxes():
    box = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=np.float64) * 0.5
    t = IDENT.copy()
    t[9:] = [3.0, 0, 0]
    assert abs(oracle.gjk(box, IDENT, box, t) - 2.0) < 1e-12              # face - face
    t[9:] = [2.0, 2.0, 0]
    assert abs(oracle.gjk(box, IDENT, box, t) - np.sqrt(2.0)) < 1e-12     # edge - edge
    t[9:] = [2.0, 2.0, 2.0]
    assert abs(oracle.gjk(box, IDENT, box, t) - np.sqrt(3.0)) < 1e-12     # vertex - vertex
    t[9:] = [0.9, 0.2, 0.1]
    assert oracle.gjk(box, IDENT, box, t) == 0.0                           # overlapping
    # rotated 45 degrees about z: corner against face
    c = np.cos(np.pi / 4)
    t = np.array([c, -c, 0, c, c, 0, 0, 0, 1, 3.0, 0, 0])
    assert abs(oracle.gjk(box, IDENT, box, t) - (3.0 - 0.5 - np.sqrt(0.5))) < 1e-12


def test_gjk_early_out_contract():
    box = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=np.float64) * 0.5
    t = IDENT.copy()
    t[9:] = [3.0, 0.3, 0.1]
    exact = oracle.gjk(box, IDENT, box, t)
    d = oracle.gjk(box, IDENT, box, t, upper=1.0)
    assert d >= 1.0 and exact >= 1.0   # contract: some value >= upper once the distance is proven >= upper


def test_gjk_matches_qp_on_scene_hulls(space_bm_scene):
    rng = np.random.default_rng(1)
    shapes = space_bm_scene.shapes
    worst, separated = 0.0, 0
    for _ in range(40):
        ia, ib = rng.integers(0, len(shapes), 2)
        va, vb = shapes[ia]["verts"], shapes[ib]["verts"]
        va = va[rng.choice(len(va), min(len(va), 40), replace=False)]
        vb = vb[rng.choice(len(vb), min(len(vb), 40), replace=False)]
        ta, tb = rand_xf(rng), rand_xf(rng)
        d = oracle.gjk(va, ta, vb, tb)
        wa = va @ ta[:9].reshape(3, 3).T + ta[9:]
        wb = vb @ tb[:9].reshape(3, 3).T + tb[9:]
        worst = max(worst, abs(d - qp_distance(wa, wb)))
        separated += d > 0
    assert worst < 1e-7 and separated > 10


def test_distance_conventions(space_scene):
    """Bullet margins: distance = core distance - 2 mm, capped at the query distance."""
    q = np.zeros(7)                                 # upright arm: nothing near the table or the planets' far side
    ob = np.zeros(16)
    ob[0] = 300
    ds, dself, dm = oracle.distances(space_scene, q, ob)
    assert ds == space_scene.struct.static_cap and dself == space_scene.struct.static_cap   # Q5: no self pairs
    assert 0 < dm <= 0.602
    ob2 = ob.copy()
    ob2[1] = 1.0                                    # latched contact short-circuits to exactly 0 (Q9)
    assert oracle.distances(space_scene, q, ob2)[2] == 0.0
    # random poses: recompose the static distance in Python from FK frames + pairwise GJK - 2 margins
    rng = np.random.default_rng(3)
    below = 0
    for _ in range(60):
        q = rng.uniform(space_scene.pos_lo, space_scene.pos_hi)
        ds, _, _ = oracle.distances(space_scene, q, ob)
        fr = oracle.fk(space_scene, q)
        best = space_scene.struct.static_cap
        for ia, ib in space_scene.static_pairs:
            a, b = space_scene.shapes[ia], space_scene.shapes[ib]
            d = oracle.gjk(a["verts"], fr[a["frame"]], b["verts"], fr[b["frame"]]) - a["margin"] - b["margin"]
            if d <= space_scene.struct.static_cap:
                best = min(best, d)
        assert abs(ds - best) < 1e-12
        below += ds < space_scene.struct.static_cap
    assert below > 5


def test_table_box_core_is_shrunk_by_the_margin(space_scene):
    table = [s for s in space_scene.shapes if s["frame"] == 0 and s["cnt"] == 8][-1]
    v = table["verts"]
    assert np.allclose(v.max(0), [0.6 - 0.001, 0.8 - 0.001, -0.001]) and np.allclose(v.min(0), [-0.599, -0.799, -0.199])
