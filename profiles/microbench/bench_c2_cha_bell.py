"""Config C2: the Cha & Bell (2024) RL shallow-water / slab boundary-layer run of
models/cha_bell2024/Oneway_ShallowWater_Slab.jl (100 cells, 181,800 points, 6 variables, ts = 3 s).
Measured on B200: 2188 timesteps/s (457 us/step, 22 launches/step) -- the 28,800-step 24 h run takes 13 s."""
import sys, time, numpy as np
from pathlib import Path; sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import scythe_jl_b200 as S
B=S.CubicBSpline
names=["h","u","v","ub","vb","wb"]
BCL={"h":B.R1T1,"u":B.R1T0,"v":B.R1T0,"ub":B.R1T0,"vb":B.R1T0,"wb":B.R1T1}
BCR={"h":B.R0,"u":B.R1T1,"v":B.R0,"ub":B.R1T1,"vb":B.R0,"wb":B.R0}
gp=S.GridParameters(geometry="RL",xmin=0,xmax=3e5,num_cells=100,BCL=BCL,BCR=BCR,vars={n:i+1 for i,n in enumerate(names)})
mp=S.ModelParameters(ts=3.0,integration_time=86400.0,output_interval=120.0,equation_set="Oneway_ShallowWater_Slab",grid_params=gp,
   physical_params=dict(g=9.81,K=5000.0,Cd=2.4e-3,Hfree=2000.0,Hb=1000.0,f=5e-5))
m=S.Model(mp,num_tiles=1)
g=m.patch
pts=S.getGridpoints(g); r,l=pts[:,0],pts[:,1]
Rmax,V0=5e4,50.0/5e4
vbar=np.where(r<Rmax,V0*r,Rmax*Rmax*V0/r)
ic=np.zeros((r.size,6)); ic[:,2]=vbar*(1+0.05*np.cos(2*l)); ic[:,4]=vbar; ic[:,3]=-0.1*vbar; ic[:,0]=100*np.exp(-(r/1e5)**2)
m.initialize(ic)
m.run(50); m.sync()
l0=m.launch_count()
t0=time.perf_counter(); m.run(2000); m.sync(); dt=time.perf_counter()-t0
print("C2 Oneway_ShallowWater_Slab: N=%d, %.1f timesteps/s, %.1f us/step, %.1f launches/step"%(g.N,2000/dt,dt/2000*1e6,(m.launch_count()-l0)/2000))
out=m.output(); print("finite", np.isfinite(out).all())
