import sys, time; sys.path.insert(0,'.')
import numpy as np
from pyscf_b200 import gto
from pyscf_b200.jk import VHFOpt
from oracle import oracle as O
from pyscf_b200.gto.mole import geometry
BENZENE = geometry("benzene")
import __graft_entry__ as g
g.smoke()
for basis in ['cc-pvdz','cc-pvtz']:
    mol = gto.M(atom=BENZENE, basis=basis)
    nao = mol.nao
    np.random.seed(1); dm = np.random.random((nao,nao)); dm = dm+dm.T
    t=time.time(); opt = VHFOpt(mol); print(basis,'setup',time.time()-t)
    for it in range(3):
        t=time.time(); vj,vk = opt.get_jk(dm, hermi=1); dt=time.time()-t
        print(basis, 'get_jk wall', dt, opt.stats())
    if basis=='cc-pvdz':
        t=time.time(); rj,rk = O.get_jk(mol, dm); print('oracle time', time.time()-t)
        print('errs', abs(vj-rj).max(), abs(vk-rk).max())
    np.save('gpurun_out/bz_%s_vj.npy'%basis, vj); np.save('gpurun_out/bz_%s_vk.npy'%basis, vk)
