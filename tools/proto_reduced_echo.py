"""Design prototype (CPU, numpy) for the REDUCED echo-space kernels (csrc/met2_t2_echo.cu, csrc/met2_basis.cu): the X2
search with its Tikhonov solves in the R-dimensional range of the dictionary (D = U C, U from an SVD here), against
voxels fitted by the unmodified reference (tests/golden/config2_subset.npz) or, for InvT2, the oracle.

    python tools/proto_reduced_echo.py n_voxels [R=24] [I|InvT2]

Measured (profiles/r02_proto_reduced_echo.txt): R = 24 and R = 20, X2-I, 300 voxels: 0 support disagreements, spectra
within 1.1e-12 / 1.2e-12, k_est within 3e-14; R = 24, X2-InvT2, 200 voxels: 0 disagreements, 9.9e-12."""
import os, sys, time
import numpy as np
from scipy.optimize import fminbound
sys.path.insert(0,'/root/repo/tools'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo')
import met2_oracle as O
from proto_dual_nnls import nnls_echo
R=int(sys.argv[2]) if len(sys.argv)>2 else 24
nvox=int(sys.argv[1])
matrix=sys.argv[3] if len(sys.argv)>3 else "I"
g=dict(np.load('/root/repo/tests/golden/config2_subset.npz'))
gr=O._grids("X2",matrix,"spline",40.0,32,10.0,1000.0)
l=np.diag(gr["L"]).copy()
sup=np.unpackbits(g["support"],axis=1)[:,:60].astype(bool); fref=np.zeros(sup.shape); fref[sup]=g["f_nz"]
sel=np.arange(0,len(g["sig"]),len(g["sig"])//nvox)[:nvox]
cache={}
def tables(a):
    if a not in cache:
        D=O.create_met2_design_matrix_epg(60,gr["T2s"],gr["T1s"],32,10.0,gr["alpha_values"][a],1000.0)
        U,s,Vt=np.linalg.svd(D,full_matrices=False)
        Ur=U[:,:R]; C=Ur.T@D
        cache[a]=(D,Ur,C)
    return cache[a]
nsup=0; wf=wk=0; 
for v in sel:
    M=g["sig"][v]; D,Ur,C=tables(int(g["fa_idx"][v]))
    b=M/M[0]
    f0,_=O.nnls(D,b); SSE=np.sum((D@f0-b)**2)
    bt=Ur.T@b; bperp=b-Ur@bt; s_perp=bperp@bperp
    Ct=C/l[None,:]
    last=[None]
    def solve(lam):
        xt=nnls_echo(Ct,bt,lam,"rec",start=last[0])
        cols=np.nonzero(xt>0)[0]; last[0]=(cols,xt[cols])
        return xt
    def obj(lam):
        xt=solve(lam); r=Ct@xt-bt
        return np.abs(r@r+s_perp-1.02*SSE)/SSE
    with np.errstate(all="ignore"):
        reg=fminbound(obj,0.0,10.0,xtol=1e-5,maxfun=300,full_output=0,disp=0)
        xt=solve(reg)
    f=xt/l*M[0]
    r=Ct@xt-bt; k=(r@r+s_perp)/SSE
    if matrix=="I": fr,kr=fref[v],g["reg"][v]
    else: fr,_s,kr=O.t2_fit_voxel(M,D,"X2",gr["L"],gr["lambda_reg"])
    nsup+= not np.array_equal(f>0,fr>0)
    wf=max(wf,np.max(np.abs(f-fr))/np.max(np.abs(fr))); wk=max(wk,abs(k-kr)/abs(kr))
print("R=%d X2-%s: %d voxels, support disagreements %d, max rel spectrum %.2e, k_est %.2e"%(R,matrix,len(sel),nsup,wf,wk))
