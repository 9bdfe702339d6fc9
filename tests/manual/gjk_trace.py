import numpy as np, torch, sys, ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from safemotionsrisk_b200 import space_backup_config, cabi
from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv
from oracle import oracle
env = SafeMotionsVecEnv(num_envs=4, config=space_backup_config(), fill_pools=False, auto_reset=False)
q=np.array([-0.6738027046181023, -1.7632363468720804, 1.1916868125458862, 1.684409065797784, 0.819326636747831, -0.05995862503146063, 1.1975051955176235])
kin=np.zeros(32); kin[:7]=q; ob=np.zeros(16)
tr=np.zeros((32,8),dtype=np.float32); res=np.zeros(128,dtype=np.float32)
lib=cabi.load()
lib.smenv_debug_gjk.argtypes=[C.c_void_p]*3+[C.c_int,C.c_int,C.c_float,C.c_void_p,C.c_void_p]
cabi.check(lib.smenv_debug_gjk(env._handle, kin.ctypes.data, ob.ctypes.data, 7, 9, C.c_float(0.102), tr.ctypes.data, res.ctypes.data),'dbg')
print('result', res[:2])
for i in range(int(res[1])+1): print(i, tr[i].tolist())
fr=oracle.fk(env.scene,q)
print('frame err', np.abs(res[4:4+96].reshape(8,12)-fr).max())
