"""Prototype checks: frame-velocity rows, arm relative velocity rows, centroidal gaps derivatives."""
import numpy as np, sys
sys.path.insert(0, '.')
from oracle.model import OracleRobot
from oracle import rbd, spatial as sp
from oracle.dynamics import *
cr=np.cross
def mxm(a,b): return np.concatenate([cr(a[3:],b[:3])+cr(a[:3],b[3:]), cr(a[3:],b[3:])])
def mxf(a,f): return np.concatenate([cr(a[3:],f[:3]), cr(a[3:],f[3:])+cr(a[:3],f[:3])])
def cstep(fun, args, idx, tangent_q=None, model=None):
    h=1e-30; x = args[idx]; n = x.size if tangent_q is None else model.nv
    cols=[]
    for d in range(n):
        e=np.zeros(n,complex); e[d]=1j*h
        a2=list(args)
        a2[idx] = rbd.integrate(model, x.astype(complex), e) if tangent_q else x+e
        cols.append(np.imag(fun(*a2))/h)
    return np.stack(cols,-1)
r = OracleRobot('b2g'); m=r.model; nv=m.nv; nb=m.njoints
rng=np.random.default_rng(5)
q=r.q0.copy(); q[7:]+=rng.normal(0,0.3,m.nq-7); qu=rng.normal(size=4); q[3:7]=qu/np.linalg.norm(qu); q[:3]=rng.normal(size=3)
v=rng.normal(size=nv); a=rng.normal(size=nv); forces=rng.normal(size=r.nf)*20
kin=rbd.Kin(m,q); R=kin.oR; p=kin.op
col_joint=[1]*6+list(range(2,nb))
def anc_j(j):
    out=[]
    while j>=1: out.append(j); j=m.parents[j]
    return out
J=np.zeros((nv,6))
for k in range(3):
    J[k]=np.concatenate([R[1][:,k],np.zeros(3)]); J[3+k]=np.concatenate([cr(p[1],R[1][:,k]),R[1][:,k]])
for j in range(2,nb):
    w=R[j]@m.axis[j]; J[m.idx_v[j]]=np.concatenate([cr(p[j],w),w])
V=[np.zeros(6) for _ in range(nb)]
for j in range(1,nb):
    V[j]=V[m.parents[j]].copy()
    for c in ([0,1,2,3,4,5] if j==1 else [m.idx_v[j]]): V[j]+=J[c]*v[c]
dyn=DynamicsWholeBodyTorque(m,r.mass,r.foot_frames)
# ---- foot velocity
fid=r.foot_frames[2]; fr=m.frames[fid]; kb=fr.parent; pk=p[kb]+R[kb]@fr.p
fv=dyn.get_frame_velocity(fid)
DQr=cstep(fv,[q,v],0,True,m)[:3]; DVr=cstep(fv,[q,v],1)[:3]
DQ=np.zeros((3,nv)); DV=np.zeros((3,nv))
for d in range(nv):
    j=col_joint[d]
    if j not in anc_j(kb): continue
    jk=J[d][:3]+cr(J[d][3:],pk); DV[:,d]=jk
    Dl=V[kb]-V[m.parents[j]]; x=mxm(J[d],Dl)
    DQ[:,d]=x[:3]+cr(x[3:],pk)+cr(V[kb][3:],jk)
print('foot vel dq', np.abs(DQ-DQr).max(), 'dv', np.abs(DV-DVr).max())
# ---- arm relative velocity
fid=r.arm_ee_frame; fr=m.frames[fid]; kb=fr.parent; pk=p[kb]+R[kb]@fr.p
fv=dyn.get_frame_velocity(fid, relative_to_base=True)
DQr=cstep(fv,[q,v],0,True,m)[:3]; DVr=cstep(fv,[q,v],1)[:3]
DQ=np.zeros((3,nv)); DV=np.zeros((3,nv)); Rb=R[1]
for d in range(nv):
    j=col_joint[d]
    if j not in anc_j(kb): continue
    jk=J[d][:3]+cr(J[d][3:],pk)
    Dl=V[kb]-V[m.parents[j]]; x=mxm(J[d],Dl)
    full = x[:3]+cr(x[3:],pk)+cr(V[kb][3:],jk)
    DV[2,d]=jk[2]; DQ[2,d]=full[2]
    if j!=1:
        DV[:2,d]=(Rb.T@jk)[:2]
        rel = x[:3]+cr(x[3:],pk)+cr(V[kb][3:]-V[1][3:],jk)
        DQ[:2,d]=(Rb.T@rel)[:2]
print('arm vel dq', np.abs(DQ-DQr).max(), 'dv', np.abs(DV-DVr).max(), np.abs(DQr).max())
# ---- centroidal acc gaps: value & derivative via composite quantities
ee=r.foot_frames+[r.ext_force_frame]
cbody=[m.frames[f].parent for f in ee]; cpos=[p[b]+R[b]@m.frames[f].p for f,b in zip(ee,cbody)]; cf=[forces[3*k:3*k+3] for k in range(len(ee))]
a0=np.array([0,0,9.81,0,0,0.]); A=[a0.copy() for _ in range(nb)]
for j in range(1,nb):
    par=m.parents[j]; A[j]=A[par].copy()
    for c in ([0,1,2,3,4,5] if j==1 else [m.idx_v[j]]): A[j]+=J[c]*a[c]+(mxm(V[par],J[c])*v[c] if j>1 else 0)
mass=m.mass; MC=[None]*nb; IB=[None]*nb; F=[None]*nb; H=[None]*nb; B22=[None]*nb
def Imul(mm_,mc,ib,mot): return np.concatenate([mm_*mot[:3]+cr(mot[3:],mc), ib@mot[3:]+cr(mc,mot[:3])])
for j in range(1,nb):
    cw=R[j]@m.com[j]+p[j]; MC[j]=mass[j]*cw; IB[j]=R[j]@m.Ic[j]@R[j].T-mass[j]*sp.skew(cw)@sp.skew(cw)
    H[j]=Imul(mass[j],MC[j],IB[j],V[j]); F[j]=Imul(mass[j],MC[j],IB[j],A[j])+mxf(V[j],H[j])
    W=sp.skew(V[j][3:]); Vx=sp.skew(V[j][:3]); C=sp.skew(MC[j])
    B22[j]=W@IB[j]-IB[j]@W-Vx@C-C@Vx-sp.skew(H[j][3:])
for k in range(len(ee)): F[cbody[k]]=F[cbody[k]]-np.concatenate([cf[k],cr(cpos[k],cf[k])])
FC=[None]+[x.copy() for x in F[1:]]; mC=list(mass); MCC=[None]+[x.copy() for x in MC[1:]]; IBC=[None]+[x.copy() for x in IB[1:]]; HC=[None]+[x.copy() for x in H[1:]]; BC=[None]+[x.copy() for x in B22[1:]]
for j in range(nb-1,1,-1):
    par=m.parents[j]; FC[par]+=FC[j]; mC[par]+=mC[j]; MCC[par]+=MCC[j]; IBC[par]+=IBC[j]; HC[par]+=HC[j]; BC[par]+=BC[j]
Mtot=mC[1]; com=MCC[1]/Mtot
def shift(Fv,c): return np.concatenate([Fv[:3], Fv[3:]-cr(c,Fv[:3])])
gaps=shift(FC[1],com)
dca=DynamicsCentroidalAcc(m,r.mass,r.foot_frames); fg=dca.dynamics_gaps(r.ext_force_frame)
print('c_acc gaps', np.abs(gaps-fg(q,v,a,forces)).max())
DQr=cstep(fg,[q,v,a,forces],0,True,m); DVr=cstep(fg,[q,v,a,forces],1); DAr=cstep(fg,[q,v,a,forces],2); DFr=cstep(fg,[q,v,a,forces],3)
DQ=np.zeros((6,nv)); DV=np.zeros((6,nv)); DA=np.zeros((6,nv)); DF=np.zeros((6,r.nf))
for d in range(nv):
    j=col_joint[d]; par=m.parents[j]
    phi=mxm(V[par],J[d]); chi=mxm(A[par],J[d])+mxm(V[par],phi); psi=mxm(V[j]+V[par],J[d])
    IC=lambda mot: Imul(mC[j],MCC[j],IBC[j],mot)
    BCm=lambda mot: np.concatenate([-2*cr(HC[j][:3],mot[3:]), BC[j]@mot[3:]])
    corr=np.zeros(6)
    for k in range(len(ee)):
        if j in anc_j(cbody[k]):
            gk=cr(J[d][3:],cf[k]); corr+=np.concatenate([gk,cr(cpos[k],gk)])
    dFq=mxf(J[d],FC[j])+IC(chi)+BCm(phi)+corr
    dc=(mC[j]*J[d][:3]+cr(J[d][3:],MCC[j]))/Mtot
    DQ[:,d]=shift(dFq,com)-np.concatenate([np.zeros(3),cr(dc,FC[1][:3])])
    DV[:,d]=shift(IC(psi)+BCm(J[d]),com); DA[:,d]=shift(IC(J[d]),com)
for k in range(len(ee)):
    for t in range(3):
        e=np.zeros(3); e[t]=1; DF[:,3*k+t]=-np.concatenate([e, cr(cpos[k]-com,e)])
for nm,X,Y in [('DQ',DQ,DQr),('DV',DV,DVr),('DA',DA,DAr),('DF',DF,DFr)]: print('c_acc',nm,np.abs(X-Y).max(),np.abs(Y).max())
# ---- centroidal vel: gaps(h,q,v)=A v - m h ; com_dyn(q,f)
dcv=DynamicsCentroidalVel(m,r.mass,r.foot_frames); hst=rng.normal(size=6)
fg=dcv.dynamics_gaps(); fc=dcv.com_dynamics(r.ext_force_frame)
print('c_vel gaps', np.abs(shift(HC[1],com)-r.mass*hst-fg(hst,q,v)).max())
DQr=cstep(fg,[hst,q,v],1,True,m); DVr=cstep(fg,[hst,q,v],2)
DQ=np.zeros((6,nv)); DV=np.zeros((6,nv))
for d in range(nv):
    j=col_joint[d]; par=m.parents[j]; phi=mxm(V[par],J[d])
    dH=mxf(J[d],HC[j])+Imul(mC[j],MCC[j],IBC[j],phi)
    dc=(mC[j]*J[d][:3]+cr(J[d][3:],MCC[j]))/Mtot
    DQ[:,d]=shift(dH,com)-np.concatenate([np.zeros(3),cr(dc,HC[1][:3])])
    DV[:,d]=shift(Imul(mC[j],MCC[j],IBC[j],J[d]),com)
print('c_vel DQ',np.abs(DQ-DQr).max(),'DV',np.abs(DV-DVr).max())
CQr=cstep(fc,[q,forces],0,True,m); CQ=np.zeros((6,nv))
for d in range(nv):
    j=col_joint[d]; dc=(mC[j]*J[d][:3]+cr(J[d][3:],MCC[j]))/Mtot
    acc=np.zeros(3)
    for k in range(len(ee)):
        dp = (J[d][:3]+cr(J[d][3:],cpos[k])) if j in anc_j(cbody[k]) else np.zeros(3)
        acc+=cr(dp-dc,cf[k])
    CQ[3:,d]=acc/r.mass
print('com_dyn DQ',np.abs(CQ-CQr).max(), np.abs(CQr).max())
