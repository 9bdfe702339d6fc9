"""First GPU sanity script: parity of the CUDA step vs the oracle on a few envs + quick timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import sys, time
import numpy as np, torch
from safemotionsrisk_b200 import space_backup_config, ball_backup_config
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv
from safemotionsrisk_b200 import abi
from oracle import oracle

def run(name, cfg, n=256, steps=20):
    env = SafeMotionsVecEnv(num_envs=n, seed=1, auto_reset=False, config=cfg)
    torch.cuda.synchronize()
    start, ball = env.pools()
    print(name, 'pool', start.shape, None if ball is None else ball.shape)
    print(' start[0] q', start[0,:7], 'v', start[0,8:15], 'ob', start[0,32:48])
    orc = oracle.OracleEnvs(env.scene, n)
    q, v, a, ob = start[:n,0:7], start[:n,8:15], start[:n,16:23], start[:n,32:48]
    env.set_state(q, v, a, ob)
    orc.set_state(q, v, a, ob)
    print(' qact diff', np.abs(env.kin.cpu().numpy()-orc.kin).max(), 'obs diff', np.abs(env.obs.cpu().numpy()-orc.obs).max())
    rng = np.random.default_rng(0)
    alive = np.ones(n, bool)
    for st in range(steps):
        act = rng.uniform(-1,1,(n,7)).astype(np.float32)
        obs, rew, done, info = env.step(act)
        torch.cuda.synchronize()
        nb = None
        if ball is not None:
            nb = np.concatenate([env.obst.cpu().numpy()[:,2:12], env.obst.cpu().numpy()[:,14:16]],axis=1)
        o_obs, o_rew, o_done, o_term, o_info = orc.step(act, nb)
        k = env.kin.cpu().numpy(); 
        dk = np.abs(k-orc.kin)[alive].max() if alive.any() else 0
        dob = np.abs(env.obst.cpu().numpy()-orc.obst)[alive].max() if alive.any() else 0
        dobs = np.abs(obs.cpu().numpy()-o_obs)[alive].max() if alive.any() else 0
        drew = np.abs(rew.cpu().numpy()-o_rew)[alive].max() if alive.any() else 0
        dinfo = np.abs(info.cpu().numpy()[:,:3]-o_info[:,:3])[alive].max() if alive.any() else 0
        dd = (done.cpu().numpy()!=o_done)[alive].sum()
        print(' step',st,'alive',alive.sum(),'kin',dk,'obst',dob,'obs',dobs,'rew',drew,'dist',dinfo,'done mism',dd, 'done',int(o_done[alive].sum()))
        alive &= (o_done==0) & (done.cpu().numpy()==0)
    env.close()

def bench(name, cfg, n=65536, steps=50):
    env = SafeMotionsVecEnv(num_envs=n, seed=1, auto_reset=True, config=cfg)
    env.reset()
    for _ in range(5): env.step_random()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): env.step_random()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/steps
    print(name, 'n', n, 'ms/step', ms, 'env-steps/s', n/ms*1e3, 'stats', env.stats.cpu().numpy()[:9])
    env.enable_counters(True); env.counters(reset=True)
    env.step_random(); c = env.counters(); print(' counters/step', {k: v/n for k,v in c.items()})
    env.close()

if __name__ == '__main__':
    t=time.time()
    run('space', space_backup_config())
    run('ball', ball_backup_config())
    run('space_bm', space_backup_config(ball_machine_mode=True), n=64, steps=10)
    print('parity time', time.time()-t)
    bench('space', space_backup_config())
    bench('ball', ball_backup_config())
