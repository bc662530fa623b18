import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import python_motionplanning_b200 as mp
from python_motionplanning_b200 import workloads as wl
np.set_printoptions(linewidth=200, precision=3)
eng = mp.Engine(0)
p = mp.VehicleParameters(); p.DFL=p.DFR=p.DRL=p.DRR=1.0
eng.set_params(p)
g = np.load('tests/golden/planar_model.npz')
sd, misc, out = eng.planar_model_batch(g["states"].T, g["torque"].T, g["mu_max"].T, g["delta"].T, g["ax_prev"], g["ay_prev"])
re = lambda a,b: np.abs(a-b)/np.maximum(np.abs(b),1.0)
print('planar sd', re(sd.cpu().numpy().T, g['state_dot']).max(axis=0))
print('planar misc', re(misc.cpu().numpy().T, g['misc']).max(axis=0))
print('planar out', re(out.cpu().numpy().T, g['outputs']).max(axis=0))
print('zero slip sd', sd.cpu().numpy()[:,1])
B,N=256,1
s0,d,t = wl.config2_rollouts(B=B,n_steps=200)
for N in (1,10,11,20,200):
    fast = eng.rollout(s0,d,t,1e-4,N,hold=10).state_end.cpu().numpy()
    z=np.zeros_like(d); d4=np.concatenate([d,d,z,z],axis=1); t4=np.repeat(t,4,axis=1)
    gen = eng.rollout(s0,d4,t4,1e-4,N,hold=10).state_end.cpu().numpy()
    mu = eng.rollout(s0,d,t,1e-4,N,hold=10,mu=np.ones((4,B))).state_end.cpu().numpy()
    ps = eng.rollout(s0,d,t,1e-4,N,hold=10,param_set=np.zeros(B,dtype=np.int32)).state_end.cpu().numpy()
    print('N',N,'gen', re(gen,fast).max(axis=1)); print('  mu', re(mu,fast).max(axis=1)); print('  ps', re(ps,fast).max(axis=1))
