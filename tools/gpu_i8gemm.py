import sys, time; sys.path.insert(0,'.')
import numpy as np
from pyscf_b200 import gto, lib as L
mol = gto.M(atom='He 0 0 0', basis='sto-3g')
h = L.Handle(mol._atm, mol._bas, mol._env)
rng = np.random.RandomState(0)
for (M,N,K,sym) in [(128,256,128,0),(300,500,1000,0),(840,840,11520,1),(2000,180,840,0)]:
    A = rng.standard_normal((M,K))*rng.uniform(0.01,3,(M,1))
    B = A if sym else rng.standard_normal((N,K))
    ref = A@B.T
    for ns in ([7] if M>128 else [1,4,7]):
        C = np.zeros((M,N))
        rc = h.lib.b200jk_i8gemm_test(h._h, M,N,K, L.dptr(A), L.dptr(B), L.dptr(C), ns, sym)
        if rc: print('ERR', h.lib.b200jk_last_error(h._h)); break
        if sym: C = np.triu(C)+np.triu(C,1).T
        ms = h.stats()['ms_kernels']
        print((M,N,K,sym), 'ns',ns,'maxerr', abs(C-ref).max(), 'ref max', abs(ref).max(), 'ms', ms, 'int8 TOPS', ns*(ns+1)/2*2*M*N*K/(ms*1e-3)/1e12*(0.5 if sym else 1))
print('--- throughput')
for (M,N,K,ns) in [(4096,4096,4096,1),(4096,4096,4096,1),(8192,8192,4096,1),(4096,4096,4096,4),(2048,2048,16384,7)]:
    A = rng.standard_normal((M,K)); B = rng.standard_normal((N,K))
    C = np.zeros((M,N))
    rc = h.lib.b200jk_i8gemm_test(h._h, M,N,K, L.dptr(A), L.dptr(B), L.dptr(C), ns, 0)
    ms = h.stats()['ms_kernels']
    print((M,N,K), 'ns',ns,'ms', ms, 'int8 TOPS', ns*(ns+1)/2*2*M*N*K/(ms*1e-3)/1e12, 'tiles', (M//128)*(N//256))
