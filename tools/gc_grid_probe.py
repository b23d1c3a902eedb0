import sys, json, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/oracle')
import helpers as H, oracle as O
from rapt_b200 import engine as eng, _lib, fields, synth, m_el, e, Re
_lib.init(0)
class SynthGrid(fields.Grid):
    def parsefile(self, fn): return synth.dipole_grid_slice(int(fn))
d,_=H.load("grid_synthetic"); files=[str(s) for s in d["files"]]
G=H.synthetic_grid(files)
fo=O.make_grid_field(G["t"],G["x"],G["y"],G["z"],G["B"],G["E"]); fg=SynthGrid(files)
traj=d["g_traj"]; mass,q,v=float(d["g_mass"]),float(d["g_charge"]),float(d["g_v"])
for arith in ("strict","fast"):
    ppar,mu=eng.gc_construct(fg,0.0,d["g_pos"],v,60.0,mass,arith=arith)
    st0=np.concatenate(([0.0],d["g_pos"],ppar))
    o=eng.gc_advance(fg,st0,mu,v,mass,q,0.05,2.6,store_every=1,max_rows=80,arith=arith)
    n=int(o["nstored"][0]); rows=o["rows"][0,:n]
    err=np.linalg.norm(rows[:,1:4]-traj[:,1:4],axis=1)/np.linalg.norm(traj[:,1:4],axis=1)
    print(arith,n,o["counters"][0],d["g_counters"].sum(0)); print(np.array2string(err[:12],precision=2))
    # tiny-step RHS probe
    rng=np.random.default_rng(3); m=256
    pos=np.column_stack([rng.uniform(4,7,m),rng.uniform(-1,1,m),rng.uniform(-1.5,1.5,m)])*Re
    pp,mu2=eng.gc_construct(fg,0.3,pos,np.full(m,v),rng.uniform(20,160,m),mass,arith=arith)
    s0=np.column_stack([np.full(m,0.3),pos,pp])
    og=eng.gc_advance(fg,s0,mu2,v,mass,q,1e-6,1e-6,arith=arith)
    ppo,muo=O.gc_construct(fo,0.3,pos,np.full(m,v),rng.uniform(20,160,m)*0+0, mass) if False else (pp,mu2)
    oo=O.gc_advance(fo,O.make_params(),s0,mu2,v,mass,q,1e-6,1e-6,store_every=0)
    dg=og["state"][:,1:5]-s0[:,1:5]; do=oo["state"][:,1:5]-s0[:,1:5]
    print(' rhs probe: pos', np.max(np.linalg.norm(dg[:,:3]-do[:,:3],axis=1)/np.linalg.norm(do[:,:3],axis=1)), 'ppar', np.max(np.abs(dg[:,3]-do[:,3]))/np.max(np.abs(do[:,3])))
