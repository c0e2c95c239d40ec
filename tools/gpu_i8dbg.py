import sys, os; sys.path.insert(0,'.')
os.environ['B200JK_I8_DEBUG']='1'
import numpy as np
from pyscf_b200 import gto, lib as L
mol = gto.M(atom='He 0 0 0', basis='sto-3g')
h = L.Handle(mol._atm, mol._bas, mol._env)
rng = np.random.RandomState(0)
for (M,N,K,ns) in [(4096,4096,4096,1),(4096,4096,4096,1),(4096,4096,4096,3),(128,256,4096,3)]:
    A = rng.standard_normal((M,K)); B = rng.standard_normal((N,K)); C=np.zeros((M,N))
    h.lib.b200jk_i8gemm_test(h._h, M,N,K, L.dptr(A), L.dptr(B), L.dptr(C), ns, 0)
    print((M,N,K,ns), h.stats()['ms_kernels'], flush=True)
