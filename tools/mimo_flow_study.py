"""Host emulation of the multi-input fast path's round loop (bmpc.cu::enqueue_round) on the MIMO side bench's workload, with the
device code compiled for the host: cold solve = [tile ADMM 25 -> polish from the iterate, all-at-once updates, cap 24] then straggler
rounds [ADMM chunk doubling the total -> polish in exchange mode, cap 24]; warm solve = [polish from the shifted sets, all-at-once,
cap 12] then straggler rounds [ADMM 100, 100, 200 -> polish from the iterate in exchange mode, cap 12].  Prints the distribution of
rounds per solve and how many instances would fall through to the Schur-form polish (after 200 iterations).
Usage: python tools/mimo_flow_study.py [instances] [warm steps]"""
import os, sys, numpy as np, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "hostemu"))
from emu import EmuSystem
from pympc_b200.workloads import mimo
cfg=mimo(); n=int(sys.argv[1]) if len(sys.argv)>1 else 100; steps=int(sys.argv[2]) if len(sys.argv)>2 else 3
rng=np.random.default_rng(4); X0=0.3*rng.standard_normal((16384,8))
def cold(E,x,um1):
    total=0; chunk=25; r=0; its=0
    while total<=200:
        E.admm_only(x,um1,cfg["xref"],chunk); total+=chunk
        vs=E.v.copy()
        if hasattr(E,"mcodes"): del E.mcodes
        U,ps=E.tpm_step(x,um1,cfg["xref"],mode=2,max_ref=24,exchange_from=-1 if r==0 else 0)
        r+=1
        if ps>0: return U,r,total
        E.v=vs; chunk=total
    return None,r,total
def warm(E,x,um1,codes,plan):
    E.mcodes,E.Uplan=codes.copy(),plan.copy(); vprev=E.v.copy()
    U,ps=E.tpm_step(x,um1,cfg["xref"],mode=1,max_ref=12)
    if ps>0: return U,1,0
    E.v=vprev; E.x=plan.copy(); E.cold=0; E.lvl=2
    total=0; chunk=100; r=1
    while total<=200:
        E.admm_only(x,um1,cfg["xref"],chunk); total+=chunk
        vs=E.v.copy()
        U,ps=E.tpm_step(x,um1,cfg["xref"],mode=2,max_ref=12,exchange_from=0)
        r+=1
        if ps>0: return U,r,total
        E.v=vs; chunk=total
    return None,r,total
t0=time.time(); rc=[]; rw=[]; fail=0; itc=[]; itw=[]
for b in range(n):
    E=EmuSystem(cfg); x=X0[b].copy(); um1=np.zeros(4)
    U,r,it=cold(E,x,um1); rc.append(r); itc.append(it)
    if U is None: fail+=1; continue
    for t in range(steps):
        codes,plan=E.mcodes.copy(),U.copy()
        um1=U[:4].copy(); x=cfg["Ad"]@x+cfg["Bd"]@um1
        U,r,it=warm(E,x,um1,codes,plan); rw.append(r); itw.append(it)
        if U is None: fail+=1; break
print("instances",n,"to-Schur-fallback",fail,"| cold rounds",np.bincount(rc),"iters mean %.0f"%np.mean(itc),"| warm rounds",np.bincount(rw),"iters mean %.1f"%np.mean(itw),"%.0f s"%(time.time()-t0))
