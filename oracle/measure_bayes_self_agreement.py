"""TEST INFRASTRUCTURE ONLY — how well does the reference's BayesReg (bayesian_interpolation.py:84-126, through the
bitwise-equal oracle port) reproduce ITSELF when the signal is perturbed by a 1e-13 relative factor?  On the 2 048
config-2 voxels of tests/golden/methods_subset.npz with reg_matrix I.  Result (profiles/r02_bayes_self_agreement.txt):
1 of 2 048 voxels moves by more than 1e-6 (lambda by 1.0e-4, spectrum by 3.6e-5); median spectrum change 7e-12.  The
tolerance of tests/test_gpu_parity.py::test_methods_subset_vs_reference for BayesReg follows from this."""
import sys, os
sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo')
import numpy as np, met2_oracle as O, multiprocessing as mp
g2=np.load('tests/golden/config2_subset.npz'); gm=np.load('tests/golden/methods_subset.npz')
sig=g2['sig'][::10]; fa=gm['fa_spline_60'].astype(int)
gr=O._grids("BayesReg","I","spline",40.0,32,10.0,1000.0)
rel=1.0+1e-13*np.random.default_rng(7).standard_normal(sig.shape)
uniq=np.unique(fa)
Dic={a:O.create_met2_design_matrix_epg(60,gr['T2s'],gr['T1s'],32,10.0,gr['alpha_values'][a],1000.0) for a in uniq}
def work(i):
    D=Dic[fa[i]]
    f0,l0=O.BayesReg_nnls(D,sig[i]/sig[i,0],gr['L'])
    s=sig[i]*rel[i]
    f1,l1=O.BayesReg_nnls(D,s/s[0],gr['L'])
    return l0,l1,np.abs(f1-f0).max()/np.abs(f0).max()
with mp.Pool(8) as p: r=np.array(p.map(work,range(len(sig))))
print('ref lambda vs fixture', np.abs(r[:,0]-gm['BayesReg_I_reg']).max())
dl=np.abs(r[:,1]-r[:,0])/r[:,0]
print('self: lambda rel >1e-6: %d, max %.2e; spectrum >1e-6: %d, >1e-5: %d max %.2e median %.2e'%((dl>1e-6).sum(),dl.max(),(r[:,2]>1e-6).sum(),(r[:,2]>1e-5).sum(),r[:,2].max(),np.median(r[:,2])))
