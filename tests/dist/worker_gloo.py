"""world_size-2 gloo worker: sharded J/K through the CPU emulation library equals the unsharded oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch.distributed as dist

from pyscf_b200 import gto
from pyscf_b200.df import DF, TaggedDM
from pyscf_b200.jk import VHFOpt
from pyscf_b200.parallel import ShardedJK
from pyscf_b200.gto.mole import make_auxmol
from oracle import oracle as O

emu = sys.argv[1]
dist.init_process_group('gloo')
rank = dist.get_rank()
mol = gto.M(atom='O 0 0 0; H 0 -0.757 0.587; H 0 0.757 0.587', basis='cc-pvdz')
nao = mol.nao
np.random.seed(1)
dm = np.random.random((nao, nao))
dm = dm + dm.T
# 4-center path
sj = ShardedJK(VHFOpt(mol, libpath=emu))
vj, vk = sj.get_jk(dm, hermi=1)
rj, rk = O.get_jk(mol, dm)
assert abs(vj - rj).max() < 1e-10 and abs(vk - rk).max() < 1e-10, (abs(vj - rj).max(), abs(vk - rk).max())
# partial results really are partial (each rank did only part of the work)
pj, pk = sj.engine.get_jk(dm, hermi=1)
assert abs(pj - rj).max() > 1e-3
# the partition on a cost table handed in by the host layer (b200jk_set_class_costs; on GPUs the table is measured by
# parallel.calibrate_partition): with trusted costs every class goes WHOLE to one rank, and the sum must not change
import ctypes
from pyscf_b200 import lib as _lib
h = sj.h
table = np.zeros(100)
for cb in range(10):
    for ck in range(cb + 1):
        table[cb * 10 + ck] = 0.05 + 0.01 * ((7 * cb + 3 * ck) % 11)     # any positive table, the same on both ranks
h.check(h.lib.b200jk_set_class_costs(h._h, _lib.dptr(table), 100), 'b200jk_set_class_costs')
vj2, vk2 = sj.get_jk(dm, hermi=1)
assert abs(vj2 - rj).max() < 1e-10 and abs(vk2 - rk).max() < 1e-10
pj2, pk2 = sj.engine.get_jk(dm, hermi=1)
assert abs(pj2 - pj).max() > 1e-6          # a different split of the work than the model-based one
vk3 = sj.get_jk(dm, hermi=1, with_j=False)[1]     # K alone: one output allocated / reduced
assert abs(vk3 - rk).max() < 1e-10
h.check(h.lib.b200jk_set_class_costs(h._h, None, 0), 'b200jk_set_class_costs')
# DF path
dfobj = DF(mol, 'weigend', libpath=emu).build()
sd = ShardedJK(dfobj)
ref, _ = O.cholesky_eri(mol, make_auxmol(mol, 'weigend'))
vj, vk = sd.get_jk(dm)
rj, rk = O.df_get_jk(ref, nao, dm)
assert abs(vj - rj).max() < 1e-10 and abs(vk - rk).max() < 1e-10
# orbital-tagged density: the occupied-orbital K engine on every rank's rows (the emulation runs the same algebra)
c = np.linalg.qr(np.random.random((nao, 5)))[0]
dmo = TaggedDM(2 * c.dot(c.T), mo_coeff=c, mo_occ=np.full(5, 2.0))
vjo, vko = sd.get_jk(dmo)
rjo, rko = O.df_get_jk(ref, nao, np.asarray(dmo))
assert abs(vjo - rjo).max() < 1e-10 and abs(vko - rko).max() < 1e-10
# sharded BUILD: each rank builds only its rows of cderi
dfs = DF(mol, 'weigend', libpath=emu, shard=(rank, dist.get_world_size())).build()
assert dfs._cderi.shape[0] < ref.shape[0]
vj, vk = ShardedJK(dfs).get_jk(dm)
assert abs(vj - rj).max() < 1e-10 and abs(vk - rk).max() < 1e-10
lo = ref.shape[0] * rank // dist.get_world_size()
assert abs(dfs._cderi - ref[lo:lo + dfs._cderi.shape[0]]).max() < 1e-9
dist.barrier()
if rank == 0:
    print('GLOO_SHARD_OK')
dist.destroy_process_group()
