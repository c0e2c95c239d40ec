import sys, time, os; sys.path.insert(0,'.'); os.environ['B200JK_DF_PROFILE']='1'
import numpy as np
from pyscf_b200 import gto
from pyscf_b200.df import DF, TaggedDM
from pyscf_b200.gto.mole import geometry
mol = gto.M(atom=geometry('c60'), basis='def2-svp'); nao=mol.nao; nocc=180
t=time.time(); d = DF(mol).build(); print('DF build s', time.time()-t, flush=True)
t=time.time(); d2 = DF(mol).build(); print('DF build (2nd) s', time.time()-t, flush=True)
rng = np.random.RandomState(1); c,_ = np.linalg.qr(rng.standard_normal((nao,nocc))); occ=np.full(nocc,2.0)
dm = TaggedDM((c*occ).dot(c.T), mo_coeff=c, mo_occ=occ)
d.get_jk(dm, with_j=False)

t=time.time(); d.get_jk(dm, with_j=False); print('K only (profiled, synced)', time.time()-t)
