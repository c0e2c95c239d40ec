"""Multi-GPU J/K: one process per GPU, work sharded inside the library, ONE all-reduce per Fock build.

Reference analogue: the OpenMP work split of CVHFnr_direct_drv with per-thread private J/K tiles reduced in a
critical section (pyscf/lib/vhf/nr_direct.c:414-482) and the additive aux-block loop of df_jk.get_jk
(pyscf/df/df_jk.py:362-380).  Here the "threads" are GPUs: shell-pair batches (4-center) or auxiliary-row
ranges (DF) are dealt to ranks by b200jk_set_shard, every rank produces partial J/K in its HBM, and the partials
are summed with a single torch.distributed all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
import ctypes

import numpy as np

from . import lib as _lib


def _dist():
    import torch.distributed as dist
    return dist


def calibrate_partition(handle, dms, rank, world, hermi=1, with_j=True, with_k=True):
    """Measure the class times of ONE unsharded, profiled 4-center build on this rank's GPU, agree on rank 0's table and hand it
    to the library as the cost table of the multi-GPU partition (b200jk_set_class_costs): with measured costs every class that
    is small against a rank's share is given whole to one rank (longest first) instead of being split eight ways.
    dms: host array [n_dm, nao, nao]; call once per (molecule, screening setup) before the sharded builds."""
    import torch
    dist = _dist()
    h = handle
    dms = np.ascontiguousarray(dms, dtype=np.float64)
    n_dm, nao = dms.shape[0], dms.shape[-1]
    vj = np.empty_like(dms) if with_j else None
    vk = np.empty_like(dms) if with_k else None
    h.check(h.lib.b200jk_set_shard(h._h, 0, 1), 'b200jk_set_shard')
    h.lib.b200jk_set_profile(h._h, 1)
    try:
        best = None
        for _ in range(2):      # the first pass pays cold caches
            h.check(h.lib.b200jk_direct_jk(h._h, _lib.dptr(dms), n_dm, nao, int(hermi), _lib.dptr(vj), _lib.dptr(vk)), 'b200jk_direct_jk')
            ms = np.zeros(100)
            h.check(h.lib.b200jk_get_class_times(h._h, _lib.dptr(ms), 100), 'b200jk_get_class_times')
            best = ms if best is None else np.minimum(best, ms)
    finally:
        h.lib.b200jk_set_profile(h._h, 0)
        h.check(h.lib.b200jk_set_shard(h._h, rank, world), 'b200jk_set_shard')
    t = torch.from_numpy(best.copy())
    if dist.is_initialized() and world > 1:
        on_gpu = torch.cuda.is_available() and dist.get_backend() == 'nccl'
        if on_gpu:
            t = t.to(torch.device('cuda', torch.cuda.current_device()))
        dist.broadcast(t, src=0)
        t = t.cpu()
    table = np.ascontiguousarray(t.numpy(), dtype=np.float64)
    h.check(h.lib.b200jk_set_class_costs(h._h, _lib.dptr(table), 100), 'b200jk_set_class_costs')
    return table


class ShardedJK:
    """Wraps a jk.VHFOpt (4-center) or df.DF (density fitting) built on THIS rank's GPU."""

    def __init__(self, engine, rank=None, world=None):
        dist = _dist()
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.engine = engine
        self.is_df = hasattr(engine, 'get_naoaux')
        h = engine._handle if self.is_df else engine.handle
        if h is None:
            engine.build()
            h = engine._handle
        self.h = h
        h.check(h.lib.b200jk_set_shard(h._h, self.rank, self.world), 'b200jk_set_shard')
        self.nao = engine.nao
        self._calibrated = self.is_df or self.world == 1     # 4-center: class times are measured at the first build

    def get_jk(self, dm, hermi=1, with_j=True, with_k=True, device_tensors=None):
        """Partial J/K on this rank, then one all-reduce.  Returns numpy arrays (every rank gets the sum)."""
        import torch
        dist = _dist()
        nao = self.nao
        dm_np = np.asarray(dm)
        shape = dm_np.shape
        dms = np.ascontiguousarray(dm_np.reshape(-1, nao, nao), dtype=np.float64)
        n_dm = len(dms)
        on_gpu = torch.cuda.is_available() and dist.get_backend() == 'nccl'
        occ = None
        nocc = 0
        if self.is_df and with_k and getattr(dm, 'mo_coeff', None) is not None and n_dm == 1:
            mo_occ = np.asarray(dm.mo_occ).ravel()
            mask = mo_occ > 0
            occ = np.ascontiguousarray(np.asarray(dm.mo_coeff).reshape(nao, -1)[:, mask] * np.sqrt(mo_occ[mask]))
            nocc = occ.shape[-1]
        h = self.h
        if not self._calibrated:
            if on_gpu:      # the CPU emulation has no class timers: it keeps the model-based partition
                calibrate_partition(h, dms, self.rank, self.world, hermi, with_j, with_k)
            self._calibrated = True
        if on_gpu:
            dev = torch.device('cuda', torch.cuda.current_device())
            d_dm = torch.from_numpy(dms).to(dev)
            # only what was asked for is allocated, reduced and copied back (get_k(omega) of a range-separated hybrid asks for K alone)
            nout = int(bool(with_j)) + int(bool(with_k))
            out = torch.zeros((nout, n_dm, nao, nao), dtype=torch.float64, device=dev)
            pj = ctypes.c_void_p(out[0].data_ptr()) if with_j else None
            pk = ctypes.c_void_p(out[nout - 1].data_ptr()) if with_k else None
            h.lib.b200jk_set_stream(h._h, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            if self.is_df:
                d_occ = torch.from_numpy(occ).to(dev) if occ is not None else None
                rc = h.lib.b200jk_df_jk_device(h._h, ctypes.c_void_p(d_dm.data_ptr()), n_dm, nao,
                                               ctypes.c_void_p(d_occ.data_ptr()) if d_occ is not None else None, nocc,
                                               int(hermi), pj, pk)
                h.check(rc, 'b200jk_df_jk_device')
            else:
                rc = h.lib.b200jk_direct_jk_device(h._h, ctypes.c_void_p(d_dm.data_ptr()), n_dm, nao, int(hermi), pj, pk)
                h.check(rc, 'b200jk_direct_jk_device')
            dist.all_reduce(out)          # the single collective of the build: (J, K requested) * n_dm * nao^2 doubles
            pin = getattr(self, '_pin_out', None)       # device -> pinned host staging (pageable D2H runs at a fraction of the link)
            if pin is None or pin.shape != out.shape:
                pin = self._pin_out = torch.empty(out.shape, dtype=torch.float64).pin_memory()
            pin.copy_(out, non_blocking=False)
            host = pin.numpy().copy()
            res = [host[0] if with_j else None, host[nout - 1] if with_k else None]
        else:
            vj = np.zeros_like(dms) if with_j else None
            vk = np.zeros_like(dms) if with_k else None
            if self.is_df:
                rc = h.lib.b200jk_df_jk(h._h, _lib.dptr(dms), n_dm, nao, _lib.dptr(occ), nocc, int(hermi), _lib.dptr(vj),
                                        _lib.dptr(vk))
                h.check(rc, 'b200jk_df_jk')
            else:
                rc = h.lib.b200jk_direct_jk(h._h, _lib.dptr(dms), n_dm, nao, int(hermi), _lib.dptr(vj), _lib.dptr(vk))
                h.check(rc, 'b200jk_direct_jk')
            buf = torch.zeros((2, n_dm, nao, nao), dtype=torch.float64)
            if with_j:
                buf[0] = torch.from_numpy(vj)
            if with_k:
                buf[1] = torch.from_numpy(vk)
            dist.all_reduce(buf)
            res = buf.numpy()
        vj = res[0].reshape(shape) if with_j else None
        vk = res[1].reshape(shape) if with_k else None
        return vj, vk
