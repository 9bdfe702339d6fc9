import numpy as np, torch, os
from oracle import oracle
from safemotionsrisk_b200 import ball_backup_config, abi
from safemotionsrisk_b200.scene import Scene
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv
g = np.load('tests/golden/ball_bm.npz')
cfg = ball_backup_config(ball_machine_mode=True)
n = g['q'].shape[0]
env = SafeMotionsVecEnv(num_envs=n, config=cfg, fill_pools=False, auto_reset=False)
sc = env.scene
# scan all golden states for self-distance mismatches
bad = []
for s in range(20):
    kin = g['out_kin'][s]; ob = g['out_obst'][s]
    ds, dse, dm = (x.cpu().numpy() for x in env.distances(kin, ob))
    ref = np.array([oracle.distances(sc, kin[e,:7], ob[e]) for e in range(n)])
    for e in range(n):
        if abs(dse[e]-ref[e,1]) > 1e-4: bad.append((s,e,dse[e],ref[e,1]))
print('mismatches', bad[:10])
if bad:
    s,e,_,_ = bad[0]
    q = g['out_kin'][s][e,:7]
    fr = oracle.fk(sc, q)
    print('q', q.tolist())
    for ia, ib in sc.self_pairs:
        a, b = sc.shapes[ia], sc.shapes[ib]
        d = oracle.gjk(a['verts'], fr[a['frame']], b['verts'], fr[b['frame']]) - 0.002
        # device single pair
        sc2 = Scene(cfg)
        sc2.struct.n_self_pairs = 1; sc2.struct.self_pairs[0][0] = ia; sc2.struct.self_pairs[0][1] = ib
        sc2.struct.n_static_pairs = 0
        e2 = SafeMotionsVecEnv.__new__(SafeMotionsVecEnv)
        import ctypes as C
        from safemotionsrisk_b200 import cabi
        lib = cabi.load(); h = C.c_void_p()
        cabi.check(lib.smenv_create(sc2.pointer(), 1, 0, 0, C.byref(h)), 'create')
        kin1 = torch.zeros((1,32), dtype=torch.float64, device='cuda'); kin1[0,:7] = torch.tensor(q)
        ob1 = torch.zeros((1,16), dtype=torch.float64, device='cuda')
        o = [torch.zeros(1, dtype=torch.float32, device='cuda') for _ in range(3)]
        cabi.check(lib.smenv_distances(h, kin1.data_ptr(), ob1.data_ptr(), o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), 1, None), 'dist')
        torch.cuda.synchronize()
        print('pair', ia, ib, 'frames', a['frame'], b['frame'], 'cnt', a['cnt'], b['cnt'], 'oracle', round(d,6), 'device', float(o[1][0]))
        lib.smenv_destroy(h)
