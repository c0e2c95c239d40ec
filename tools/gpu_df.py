import sys, time, os; sys.path.insert(0,'.')
import numpy as np
from pyscf_b200 import gto
from pyscf_b200.df import DF, TaggedDM
from pyscf_b200.gto.mole import geometry
name = sys.argv[1] if len(sys.argv)>1 else 'c60'
basis = sys.argv[2] if len(sys.argv)>2 else 'def2-svp'
nocc = int(sys.argv[3]) if len(sys.argv)>3 else 180
mol = gto.M(atom=geometry(name), basis=basis)
nao = mol.nao
t=time.time(); d = DF(mol).build(); print('DF build s', time.time()-t, 'naux', d.get_naoaux(), 'nao', nao, flush=True)
rng = np.random.RandomState(1)
c,_ = np.linalg.qr(rng.standard_normal((nao,nocc)))
occ = np.full(nocc, 2.0)
dm = TaggedDM((c*occ).dot(c.T), mo_coeff=c, mo_occ=occ)
res = {}
for eng in ['dgemm','tcgen05']:
    d.set_k_engine(eng, 7)
    for it in range(3):
        t=time.time(); vj,vk = d.get_jk(dm); dt=time.time()-t
    print(eng, 'get_jk s', dt, 'kernels ms', d.stats()['ms_kernels'], flush=True)
    res[eng]=vk
print('tcgen05 vs dgemm  max|dK|', abs(res['tcgen05']-res['dgemm']).max(), ' max|K|', abs(res['dgemm']).max())
for ns in [5,6,8]:
    d.set_k_engine('tcgen05', ns); _,vk = d.get_jk(dm); print('ns',ns,'err', abs(vk-res['dgemm']).max(), d.stats()['ms_kernels'])
t=time.time(); vj3,_ = d.get_jk(dm, with_k=False); print('J only s', time.time()-t, d.stats()['ms_kernels'])
