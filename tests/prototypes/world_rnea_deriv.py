"""Prototype (numpy) of the world-frame, column-per-lane RNEA + analytic derivative scheme used by the
CUDA node kernels.  Checked against the oracle's complex-step Jacobians.  Design aid, not product."""
import numpy as np, sys
sys.path.insert(0, '.')
from oracle.model import OracleRobot
from oracle import rbd, spatial as sp
from oracle.dynamics import DynamicsWholeBodyTorque

def cr(a,b): return np.cross(a,b)
def mxm(a,b): return np.concatenate([cr(a[3:],b[:3])+cr(a[:3],b[3:]), cr(a[3:],b[3:])])
def mxf(a,f): return np.concatenate([cr(a[3:],f[:3]), cr(a[3:],f[3:])+cr(a[:3],f[:3])])

def run(name='b2g', seed=0):
    r = OracleRobot(name); m = r.model; nv = m.nv
    rng = np.random.default_rng(seed)
    q = r.q0.copy(); q[7:] += rng.normal(0,0.3,m.nq-7); quat = rng.normal(size=4); q[3:7]=quat/np.linalg.norm(quat); q[:3]=rng.normal(size=3)
    v = rng.normal(size=nv); a = rng.normal(size=nv); forces = rng.normal(size=r.nf)*20
    ee = r.foot_frames + ([r.ext_force_frame] if r.ext_force_frame else [])
    kin = rbd.Kin(m, q)
    nb = m.njoints
    # column tables
    col_joint = [1]*6 + list(range(2, nb))
    def anc_cols(j):  # columns of joints that are ancestors-or-self of joint j
        cols=[]; 
        while j>=1:
            cols = ([0,1,2,3,4,5] if j==1 else [m.idx_v[j]]) + cols; j = m.parents[j]
        return cols
    a0 = np.array([0,0,9.81,0,0,0.])
    R = kin.oR; p = kin.op
    J = np.zeros((nv,6))
    for k in range(3):
        J[k] = np.concatenate([R[1][:,k], np.zeros(3)])
        J[3+k] = np.concatenate([cr(p[1], R[1][:,k]), R[1][:,k]])
    for j in range(2, nb):
        w = R[j] @ m.axis[j]; J[m.idx_v[j]] = np.concatenate([cr(p[j], w), w])
    V = [np.zeros(6)]*nb; A = [None]*nb; A[0]=a0
    V = [np.zeros(6) for _ in range(nb)]; A = [a0.copy() for _ in range(nb)]
    for j in range(1, nb):
        par = m.parents[j]
        cols = [0,1,2,3,4,5] if j==1 else [m.idx_v[j]]
        V[j] = V[par].copy(); A[j] = A[par].copy()
        for c in cols:
            A[j] += J[c]*a[c] + mxm(V[par], J[c])*v[c] if j>1 else J[c]*a[c]
            V[j] += J[c]*v[c]
    # world inertias
    mass = m.mass; MC=[None]*nb; IB=[None]*nb
    def Imul(j, mot):
        vv, ww = mot[:3], mot[3:]
        return np.concatenate([mass[j]*vv + cr(ww, MC[j]), IB[j]@ww + cr(MC[j], vv)])
    F=[np.zeros(6) for _ in range(nb)]; Pl=[None]*nb; B22=[None]*nb
    cpos = {}; cforce={}
    for idx,fid in enumerate(ee):
        fr = m.frames[fid]; cpos[idx] = p[fr.parent] + R[fr.parent]@fr.p; cforce[idx]=forces[3*idx:3*idx+3]
    cbody = [m.frames[fid].parent for fid in ee]
    for j in range(1, nb):
        cw = R[j]@m.com[j] + p[j]; MC[j] = mass[j]*cw
        IB[j] = R[j]@m.Ic[j]@R[j].T - mass[j]*sp.skew(cw)@sp.skew(cw)
        h = Imul(j, V[j])
        F[j] = Imul(j, A[j]) + mxf(V[j], h)
        W = sp.skew(V[j][3:]); Vx = sp.skew(V[j][:3]); C = sp.skew(MC[j])
        Pl[j] = h[:3].copy()
        B22[j] = W@IB[j] - IB[j]@W - Vx@C - C@Vx - sp.skew(h[3:])
    for idx in range(len(ee)):
        F[cbody[idx]] = F[cbody[idx]] - np.concatenate([cforce[idx], cr(cpos[idx], cforce[idx])])
    # composites
    FC=[f.copy() for f in F]; mC=list(mass); MCC=[None]+[x.copy() for x in MC[1:]]; IBC=[None]+[x.copy() for x in IB[1:]]
    PC=[None]+[x.copy() for x in Pl[1:]]; BC=[None]+[x.copy() for x in B22[1:]]
    for j in range(nb-1, 1, -1):
        par = m.parents[j]
        FC[par]+=FC[j]; mC[par]+=mC[j]; MCC[par]+=MCC[j]; IBC[par]+=IBC[j]; PC[par]+=PC[j]; BC[par]+=BC[j]
    def ICmul(j, mot):
        vv, ww = mot[:3], mot[3:]
        return np.concatenate([mC[j]*vv + cr(ww, MCC[j]), IBC[j]@ww + cr(MCC[j], vv)])
    def BCmul(j, mot):  # B^C m
        al = mot[3:]
        return np.concatenate([-2*cr(PC[j], al), BC[j]@al])
    def BCTmul(j, mot):  # B^C^T m
        return np.concatenate([np.zeros(3), 2*cr(PC[j], mot[:3]) + BC[j].T@mot[3:]])
    tau = np.array([J[c]@FC[col_joint[c]] for c in range(nv)])
    dyn = DynamicsWholeBodyTorque(m, r.mass, r.foot_frames)
    tau_ref = dyn.rnea_dynamics(r.ext_force_frame)(q, v, a, forces)
    print('tau err', np.abs(tau - tau_ref).max())
    # per column vectors
    phi = np.zeros((nv,6)); chi=np.zeros((nv,6)); psi=np.zeros((nv,6))
    for d in range(nv):
        j = col_joint[d]; par = m.parents[j]
        phi[d] = mxm(V[par], J[d]); chi[d] = mxm(A[par], J[d]) + mxm(V[par], phi[d]); psi[d] = mxm(V[j]+V[par], J[d])
    is_anc = lambda c, d: col_joint[c] in _anc_joints(col_joint[d])
    def _anc_joints(j):
        out=[]
        while j>=1: out.append(j); j=m.parents[j]
        return out
    M = np.zeros((nv,nv)); DV=np.zeros((nv,nv)); DQ=np.zeros((nv,nv)); DF = np.zeros((nv, r.nf))
    for d in range(nv):
        j = col_joint[d]
        w_d = ICmul(j, J[d]); y_d = BCTmul(j, J[d])
        dFa = w_d
        dFv = ICmul(j, psi[d]) + BCmul(j, J[d])
        corr = np.zeros(6)
        for k in range(len(ee)):
            if j in _anc_joints(cbody[k]):
                gk = cr(J[d][3:], cforce[k]); corr += np.concatenate([gk, cr(cpos[k], gk)])
        dFq_norigid = ICmul(j, chi[d]) + BCmul(j, phi[d]) + corr
        dFq = mxf(J[d], FC[j]) + dFq_norigid
        for c in anc_cols(j):
            M[c,d] = J[c]@dFa; M[d,c]=M[c,d]
            DV[c,d] = J[c]@dFv
            same = col_joint[c]==j
            DQ[c,d] = J[c]@(dFq_norigid if same else dFq)
            if not same:  # row d (descendant), column c (strict ancestor)
                DV[d,c] = w_d@psi[c] + y_d@J[c]
                val = w_d@chi[c] + y_d@phi[c]
                for k in range(len(ee)):
                    if j in _anc_joints(cbody[k]):
                        gk = cr(J[c][3:], cforce[k]); val += J[d]@np.concatenate([gk, cr(cpos[k], gk)])
                DQ[d,c] = val
        for k in range(len(ee)):
            if j in _anc_joints(cbody[k]):
                DF[d, 3*k:3*k+3] = -(J[d][:3] + cr(J[d][3:], cpos[k]))
    # reference by complex step wrt local tangent of q
    h=1e-30; f = dyn.rnea_dynamics(r.ext_force_frame)
    DQr=np.zeros((nv,nv)); DVr=np.zeros((nv,nv)); Mr=np.zeros((nv,nv)); DFr=np.zeros((nv,r.nf))
    for d in range(nv):
        e = np.zeros(nv, complex); e[d]=1j*h
        DQr[:,d] = np.imag(f(rbd.integrate(m, q.astype(complex), e), v, a, forces))/h
        DVr[:,d] = np.imag(f(q, v+e, a, forces))/h
        Mr[:,d] = np.imag(f(q, v, a+e, forces))/h
    for d in range(r.nf):
        e = np.zeros(r.nf, complex); e[d]=1j*h
        DFr[:,d] = np.imag(f(q, v, a, forces+e))/h
    for nm, X, Y in [('M',M,Mr),('DV',DV,DVr),('DQ',DQ,DQr),('DF',DF,DFr)]:
        print(nm, 'err', np.abs(X-Y).max(), 'scale', np.abs(Y).max())
run('b2g'); run('go2', 3)
