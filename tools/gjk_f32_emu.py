"""NumPy float32 emulation of the device GJK (smenv_geom.cuh: pair_distance) for debugging on a CPU-only box."""
import numpy as np
f = np.float32

def dot(a, b): return f(a[0]*b[0] + a[1]*b[1] + a[2]*b[2])
def cross(a, b): return np.array([a[1]*b[2]-a[2]*b[1], a[2]*b[0]-a[0]*b[2], a[0]*b[1]-a[1]*b[0]], dtype=f)

def closest_segment(a, b):
    ab = b - a; t = -dot(a, ab); den = dot(ab, ab)
    if t <= 0 or not den > 0: return a.copy(), 1
    if t >= den: return b.copy(), 2
    return (a + f(t/den)*ab).astype(f), 3

def closest_triangle(a, b, c):
    ab, ac = b-a, c-a
    d1, d2 = -dot(ab, a), -dot(ac, a)
    if d1 <= 0 and d2 <= 0: return a.copy(), 1
    d3, d4 = -dot(ab, b), -dot(ac, b)
    if d3 >= 0 and d4 <= d3: return b.copy(), 2
    vc = f(d1*d4 - d3*d2)
    if vc <= 0 and d1 >= 0 and d3 <= 0 and d1-d3 > 0: return (a + f(d1/(d1-d3))*ab).astype(f), 3
    d5, d6 = -dot(ab, c), -dot(ac, c)
    if d6 >= 0 and d5 <= d6: return c.copy(), 4
    vb = f(d5*d2 - d1*d6)
    if vb <= 0 and d2 >= 0 and d6 <= 0 and d2-d6 > 0: return (a + f(d2/(d2-d6))*ac).astype(f), 5
    va = f(d3*d6 - d5*d4); e1 = f(d4-d3); e2 = f(d5-d6)
    if va <= 0 and e1 >= 0 and e2 >= 0 and e1+e2 > 0: return (b + f(e1/(e1+e2))*(c-b)).astype(f), 6
    s = f(va+vb+vc); n = cross(ab, ac)
    if s > 0 and dot(n, n) > f(1e-10)*dot(ab, ab)*dot(ac, ac):
        den = f(1)/s
        return (a + f(vb*den)*ab + f(vc*den)*ac).astype(f), 7
    cands = [closest_segment(a, b), closest_segment(a, c), closest_segment(b, c)]
    q = [dot(p, p) for p, _ in cands]
    if q[0] <= q[1] and q[0] <= q[2]: return cands[0]
    if q[1] <= q[2]:
        p, m = cands[1]; return p, (m & 1) | ((m & 2) << 1)
    p, m = cands[2]; return p, m << 1

def pair_distance(vA, TA, vB, TB, cA, cB, margin, upper=0.0, touch=-1.0, log=None):
    """vA, vB float32 [n,3]; TA, TB = (R[3,3], t[3]) float32; returns distance - margin like the device."""
    RA, tA = TA; RB, tB = TB
    if upper > 0: upper = f(upper + margin)
    if touch >= 0: touch = f(touch + margin)
    S = []; ids = []
    v = (RA @ cA + tA - (RB @ cB + tB)).astype(f); vv = dot(v, v)
    have = False
    for it in range(32):
        dA = (RA.T @ (-v)).astype(f); dB = (RB.T @ v).astype(f)
        sa = int(np.argmax(vA @ dA)); sb = int(np.argmax(vB @ dB))
        w = ((RA @ vA[sa] + tA) - (RB @ vB[sb] + tB)).astype(f); idw = (sa, sb)
        if not have:
            S = [w]; ids = [idw]; v = w; vv = dot(v, v); have = True
            continue
        vw = dot(v, w)
        if log is not None: log.append((it, len(S), float(np.sqrt(vv)), float(vw/np.sqrt(vv))))
        if upper > 0 and vw > 0 and vw*vw >= upper*upper*vv:
            vv = max(vv, upper*upper); reason = 'upper'; break
        nv = f(np.sqrt(vv))
        if vv - vw <= max(f(1e-6)*vv, f(3e-7)*nv): reason = 'converged'; break
        if idw in ids: reason = 'duplicate'; break
        S.append(w); ids.append(idw)
        if len(S) == 2:
            p, m = closest_segment(S[0], S[1]); keep = [i for i in range(2) if m & (1 << i)]
        elif len(S) == 3:
            p, m = closest_triangle(S[0], S[1], S[2]); keep = [i for i in range(3) if m & (1 << i)]
        else:
            faces = [(0, 1, 2, 3), (0, 1, 3, 2), (0, 2, 3, 1), (1, 2, 3, 0)]
            best = None; inside_all = True
            for (i, j, k, l) in faces:
                a, b, c, d = S[i], S[j], S[k], S[l]
                n = cross(b-a, c-a); sd = dot(d-a, n); so = -dot(a, n)
                if not so*sd > 0: inside_all = False
                pp, mm = closest_triangle(a, b, c); dd = dot(pp, pp)
                if best is None or dd < best[0]: best = (dd, pp, mm, (i, j, k))
            if inside_all:
                e1, e2, e3 = S[1]-S[0], S[2]-S[0], S[3]-S[0]
                det = dot(e3, cross(e1, e2))
                if det*det > f(1e-8)*dot(e1, e1)*dot(e2, e2)*dot(e3, e3):
                    vv = f(0); reason = 'enclosed'; break
            _, p, m, tri = best; keep = [tri[i] for i in range(3) if m & (1 << i)]
        S = [S[i] for i in keep]; ids = [ids[i] for i in keep]
        nd = dot(p, p)
        if not nd < vv: reason = 'noprogress'; break
        v = p.astype(f); vv = nd
        if vv <= 1e-20: vv = f(0); reason = 'zero'; break
        if touch >= 0 and vv <= touch*touch: reason = 'touch'; break
    else:
        reason = 'maxiter'
    return float(np.sqrt(vv)) - margin, reason, it
