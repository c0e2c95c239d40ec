#!/usr/bin/env python
"""bench.py — J/K Fock-build seconds per SCF iteration (BASELINE.json metric) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" = one J/K Fock build (one get_jk-equivalent call) for the workload's density matrix.
Default workload (N=1): configs[1] of BASELINE.json, benzene / cc-pVTZ RHF, 4-center direct J/K.

Timed numbers
  value / ms_per_step : J/K build with D, J, K resident in HBM (b200jk_direct_jk_device on torch's current
                        stream), CUDA events per step, 256 MiB L2 flush between steps outside the event pairs.
  e2e                 : the same build through the public plugin call VHFOpt.get_jk with pinned HOST buffers
                        (H2D of D and D2H of J,K inside the timed region).
  roofline            : direct: the class kernels (one template, 55 instantiations) against the measured FP64
                        FMA-pipe peak (b200jk_fp64_peak micro-benchmark; the 4-center path is FP64-bound,
                        SURVEY.md §8d) and, beside it, the HBM figure (algorithmic bytes / time).
                        DF: the dominant kernel (i8gemm_ar_kernel, stage 1 of DF-K) per launch, from CUDA events the
                        library records around every launch of the timed steps (b200jk_df_stage_times), against
                        2 x the measured bf16 tensor peak; the other stages (stage 2, slicing, the two HBM-bound DF-J
                        passes) and the whole-build figure are listed under roofline.stages / roofline.whole_build.
  cpu_baseline        : the CPU oracle (McMurchie-Davidson port of the reference path, OpenMP, all host cores)
                        on the same workload; rank 0, N=1 only.
--impl reference times that CPU arm alone with the same JSON schema.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    'benzene-ccpvtz-direct': dict(geom='benzene', basis='cc-pvtz', nocc=21, kind='direct'),
    'benzene-ccpvdz-direct': dict(geom='benzene', basis='cc-pvdz', nocc=21, kind='direct'),
    'h2o-ccpvdz-direct': dict(geom='h2o', basis='cc-pvdz', nocc=5, kind='direct'),
    'c60-def2svp-df': dict(geom='c60', basis='def2-svp', nocc=180, kind='df'),
    'benzene-def2svp-df': dict(geom='benzene', basis='def2-svp', nocc=21, kind='df'),
    'gly30-ccpvdz-df': dict(geom='gly30', basis='cc-pvdz', nocc=455, kind='df'),   # BASELINE config 5 (full-range J/K part)
    'taxol-def2tzvp-df': dict(geom='taxol', basis='def2-tzvp', nocc=226, kind='df'),  # BASELINE config 4 (111 GB tensor: needs >= 2 GPUs)
    'gly4-ccpvdz-df': dict(geom='gly4', basis='cc-pvdz', nocc=65, kind='df'),
}

def scf_like_dm(nao, nocc, seed=1):
    rng = np.random.RandomState(seed)
    c, _ = np.linalg.qr(rng.standard_normal((nao, nocc)))
    return 2.0 * c.dot(c.T)


def build_mol(w):
    from pyscf_b200 import gto
    from pyscf_b200.gto.mole import geometry
    return gto.M(atom=geometry(w['geom']), basis=w['basis'])


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([t.strip() for t in line.split(',')])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], s[2:6]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                continue
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx or None, 'reasons': sorted(reasons),
                'samples': len(sm)}


def algorithmic_bytes(opt):
    """Algorithmic HBM bytes of one direct build: D, J and K once each (3 n^2 doubles) plus the shell-pair
    records the kernels stream (48 B of each 64-byte record are payload)."""
    st = opt.stats()
    n = st['n_sph']
    return 3 * n * n * 8 + st['n_pairs'] * 48


def run_ours(args, rank, world):
    import torch
    import ctypes
    from pyscf_b200.jk import VHFOpt
    from pyscf_b200.df import DF, TaggedDM
    w = WORKLOADS[args.workload]
    is_df = w['kind'] == 'df'
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    mol = build_mol(w)
    nao = mol.nao
    rng = np.random.RandomState(1)
    c_occ, _ = np.linalg.qr(rng.standard_normal((nao, w['nocc'])))
    dm_h = 2.0 * c_occ.dot(c_occ.T)
    occ_h = np.ascontiguousarray(c_occ * np.sqrt(2.0))
    t0 = time.time()
    if is_df:
        eng = DF(mol, device=local, shard=(rank, world) if world > 1 else None).build()
        h = eng._handle
    else:
        eng = VHFOpt(mol, direct_scf_tol=1e-13, device=local)
        h = eng.handle
    setup_s = time.time() - t0
    # strong scaling: ONE Fock build is split over the ranks (shell-pair batches / auxiliary rows), partial J,K
    # are summed by a single NCCL all-reduce per build
    h.check(h.lib.b200jk_set_shard(h._h, rank, world), 'b200jk_set_shard')
    stream = torch.cuda.current_stream(dev)
    h.lib.b200jk_set_stream(h._h, ctypes.c_void_p(stream.cuda_stream))

    dm_d = torch.from_numpy(dm_h).to(dev)
    occ_d = torch.from_numpy(occ_h).to(dev)
    out_d = torch.zeros((2, nao, nao), dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)

    def step_device():
        if is_df:
            rc = h.lib.b200jk_df_jk_device(h._h, ctypes.c_void_p(dm_d.data_ptr()), 1, nao, ctypes.c_void_p(occ_d.data_ptr()),
                                           w['nocc'], 1, ctypes.c_void_p(out_d[0].data_ptr()), ctypes.c_void_p(out_d[1].data_ptr()))
            h.check(rc, 'b200jk_df_jk_device')
        else:
            rc = h.lib.b200jk_direct_jk_device(h._h, ctypes.c_void_p(dm_d.data_ptr()), 1, nao, 1,
                                               ctypes.c_void_p(out_d[0].data_ptr()), ctypes.c_void_p(out_d[1].data_ptr()))
            h.check(rc, 'b200jk_direct_jk_device')
        if world > 1:
            dist.all_reduce(out_d)

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms = []
    stage_log = []
    launches = 0
    torch.cuda.synchronize()
    t_wall0 = time.time()
    for k in range(args.steps):
        flush.zero_()
        if world > 1:
            dist.barrier()
        evs[k][0].record(stream)
        step_device()
        evs[k][1].record(stream)
        torch.cuda.synchronize()
        st = h.stats()
        kern_ms.append(st['ms_kernels'])
        launches += st['kernel_launches']
        if is_df:
            stage_log.append(h.df_stage_times())
    torch.cuda.synchronize()
    t_wall = time.time() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    ms_per_step = float(np.mean(step_ms))
    vj_dev = out_d[0].cpu().numpy().copy()
    vk_dev = out_d[1].cpu().numpy().copy()
    # ---- end-to-end through the public plugin call with pinned host buffers (H2D + D2H inside the timed region)
    dm_pin = torch.from_numpy(dm_h).pin_memory().numpy()
    if is_df:
        dm_pub = TaggedDM(dm_pin, mo_coeff=c_occ, mo_occ=np.full(w['nocc'], 2.0))
    else:
        dm_pub = dm_pin
    h.lib.b200jk_set_stream(h._h, None)

    def step_public():
        if world > 1:
            from pyscf_b200.parallel import ShardedJK
            if not hasattr(step_public, 'sj'):
                step_public.sj = ShardedJK(eng, rank, world)
            return step_public.sj.get_jk(dm_pub, hermi=1)
        return eng.get_jk(dm_pub, hermi=1)

    for _ in range(2):
        step_public()
    e2e_ms = []
    for k in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = time.perf_counter()
        vj, vk = step_public()
        e2e_ms.append((time.perf_counter() - t) * 1e3)
    clocks = sampler.finish()
    e2e_ms_mean = float(np.mean(e2e_ms))

    if world > 1:
        tt = torch.tensor([ms_per_step, e2e_ms_mean], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_per_step, e2e_ms_mean = float(tt[0]), float(tt[1])

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    # ---- roofline
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    kernel_ms = float(np.mean(kern_ms))
    if not is_df:
        from pyscf_b200.flops import direct_jk_flops
        peak = ctypes.c_double(0)
        h.lib.b200jk_fp64_peak(h._h, ctypes.byref(peak))
        flops, n_eri = direct_jk_flops(mol)
        bytes_alg = algorithmic_bytes(eng)
        fp64_ach = flops / world / (kernel_ms * 1e-3) / 1e12
        roof = {'bound': 'fp64', 'achieved': fp64_ach, 'peak': peak.value, 'unit': 'TFLOP/s',
                'frac': fp64_ach / peak.value if peak.value else None, 'traffic': None,
                'kernel': 'jk_class_kernel / jk_tpq_kernel <QClass<LI,LJ,LK,LL,NP>> (all class launches of one build)',
                'kernel_ms_per_step': kernel_ms, 'alg_flops_per_step': flops, 'alg_cart_eris_per_step': n_eri,
                'peak_source': 'b200jk_fp64_peak DFMA micro-benchmark (MEASURED_PEAKS.json has no fp64 entry)',
                'hbm': {'bound': 'hbm', 'achieved': bytes_alg / (kernel_ms * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                        'frac': bytes_alg / (kernel_ms * 1e-3) / 1e9 / hbm_peak, 'alg_bytes_per_step': bytes_alg,
                        'peak_source': 'MEASURED_PEAKS.json hbm_gbs (of measured)' if peaks else 'fallback 6650'}}
        path = '4-center direct J/K (hermi=1, with_j, with_k)'
    else:
        naux = eng.get_naoaux()
        ns = eng.k_slices
        nsl = ns * (ns + 1) // 2                                   # slice GEMMs actually executed (k + l < ns)
        fp64_flops = 4.0 * naux * nao * nao * w['nocc']           # dsymm + dgemm count of the reference
        int8_ops = fp64_flops * nsl
        bf16_peak = peaks.get('bf16_tflops_sustained', 1400.0)
        tensor_peak = 2 * bf16_peak
        cderi_bytes = naux * nao * (nao + 1) / 2 * 8
        # per-stage device times of the timed steps (CUDA events around every launch, b200jk_df_stage_times)
        stg = {k: (float(np.mean([t[k][0] for t in stage_log])), int(stage_log[0][k][1])) for k in stage_log[0]}
        half_ops = int8_ops / 2 / world                            # each GEMM stage carries half of the 4*naux*nao^2*nocc count
        stages = {}
        for k in ('k_gemm1', 'k_gemm2'):
            ms_k, n_k = stg[k]
            if n_k:
                stages[k] = {'kernel': 'i8gemm_ar_kernel' if k == 'k_gemm1' else 'i8gemm_kernel', 'bound': 'tensor',
                             'launches_per_step': n_k, 'ms_per_launch': ms_k / n_k, 'ms_per_step': ms_k,
                             'alg_int8_ops_per_launch': half_ops / n_k, 'achieved': half_ops / (ms_k * 1e-3) / 1e12,
                             'peak': tensor_peak, 'unit': 'TOP/s (int8)', 'frac': half_ops / (ms_k * 1e-3) / 1e12 / tensor_peak}
                if k == 'k_gemm2':
                    stages[k]['note'] = ('algorithmic count = the full Y Y^T product of the reference dgemm (SURVEY 8d); the kernel executes only '
                                         'the upper-triangle tiles, so this fraction is above the tensor-pipe activity ncu reports (69 %)')
        for k in ('j_rho', 'j_acc'):
            ms_k, n_k = stg[k]
            if n_k:
                stages[k] = {'kernel': 'dfj_rho_kernel' if k == 'j_rho' else 'dfj_acc_kernel', 'bound': 'hbm',
                             'launches_per_step': n_k, 'ms_per_step': ms_k, 'alg_bytes_per_step': cderi_bytes / world,
                             'achieved': cderi_bytes / world / (ms_k * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                             'frac': cderi_bytes / world / (ms_k * 1e-3) / 1e9 / hbm_peak}
        ms_k, n_k = stg['k_slice']
        stages['k_slice'] = {'kernel': 'rowmax_kernel + split_long_kernel (int8 slicing of Y)', 'ms_per_step': ms_k,
                             'launches_per_step': n_k}
        ach = int8_ops / world / (kernel_ms * 1e-3) / 1e12
        g1 = stages.get('k_gemm1')
        if g1:     # the dominant kernel: stage 1 of DF-K
            roof = {'bound': 'tensor', 'achieved': g1['achieved'], 'peak': tensor_peak, 'unit': 'TOP/s (int8)', 'frac': g1['frac'],
                    # dram__bytes_read.sum + dram__bytes_write.sum of one i8gemm_ar launch of this workload, ncu --set full
                    # (profiles/r01_ncu_i8ar.txt); null for other workloads
                    'traffic': 2.018729e9 + 446.094080e6 if (args.workload == 'c60-def2svp-df' and world == 1) else None,
                    'kernel': 'i8gemm_ar_kernel (tcgen05.mma.kind::i8, stage 1 of DF-K: Y = (P|mu nu) C~), CUDA events around '
                              'each of its launches inside the timed steps',
                    'ms_per_launch': g1['ms_per_launch'], 'launches_per_step': g1['launches_per_step'],
                    'alg_int8_ops_per_launch': g1['alg_int8_ops_per_launch']}
        else:      # cuBLAS DGEMM yardstick engine or general-density path: whole build only
            roof = {'bound': 'tensor', 'achieved': ach, 'peak': tensor_peak, 'unit': 'TOP/s (int8)', 'frac': ach / tensor_peak,
                    'traffic': None, 'kernel': 'whole DF J+K build'}
        roof.update({
                'slice_gemms': nsl, 'stages': stages,
                'whole_build': {'ms_per_step': kernel_ms, 'achieved': ach, 'frac': ach / tensor_peak, 'unit': 'TOP/s (int8)',
                                'note': 'all int8 slice-GEMM work over the whole DF J+K build time (J passes, slicing, both GEMM stages)',
                                'fp64_equiv_flops_per_step': fp64_flops,
                                'fp64_equiv_tflops': fp64_flops / world / (kernel_ms * 1e-3) / 1e12},
                'peak_source': 'tensor: 2 x MEASURED_PEAKS.json bf16_tflops_sustained (int8 dense = 2x bf16 on sm_100a; sustained '
                               'because the kernel runs inside a long step); hbm: MEASURED_PEAKS.json hbm_gbs'
                               if peaks else 'fallback 2 x 1400 TFLOP/s, 6650 GB/s (B200_PROFILING.md)'})
        path = 'DF J/K (cderi resident, K via tcgen05 int8 slices, %d slices)' % ns
    # ---- CPU baseline (oracle port), rank 0, N=1 only
    cpu = None
    if world == 1 and not args.no_cpu:
        if is_df:
            cpu = cpu_baseline_df(mol, dm_h, c_occ, args.workload)
        else:
            cpu = cpu_baseline(mol, dm_h, args.workload)
        if cpu.get('vj') is not None:
            cpu['max_abs_dJ_vs_gpu'] = float(abs(vj - cpu.pop('vj')).max())
            cpu['max_abs_dK_vs_gpu'] = float(abs(vk - cpu.pop('vk')).max())
    out = {
        'metric': 'J/K Fock-build wall-s/iter', 'value': ms_per_step * 1e-3, 'unit': 's',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'higher_is_better': False, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': args.workload, 'molecule': w['geom'], 'basis': w['basis'], 'nao': nao, 'path': path,
                   'direct_scf_tol': 1e-13, 'dm': 'SCF-like 2*C_occ*C_occ^T, orthonormal random C_occ, seed 1',
                   'l2_flush': '256 MiB memset between steps, outside the per-step CUDA-event pairs',
                   'parallelism': ('one build sharded over %d GPUs + 1 NCCL all-reduce of [J;K]' % world) if world > 1 else 'single GPU'},
        'e2e': {'value': e2e_ms_mean * 1e-3, 'unit': 's', 'h2d_bytes_per_step': int(nao * nao * 8 + (nao * w['nocc'] * 8 if is_df else 0)),
                'd2h_bytes_per_step': int(2 * nao * nao * 8),
                'api': 'pyscf_b200.df.DF.get_jk' if is_df else 'pyscf_b200.jk.VHFOpt.get_jk (pinned host dm)'},
        'gpu_launches': int(launches), 'setup_s': setup_s, 'clocks': clocks, 'roofline': roof,
        'wall_s_timed_region': t_wall,
        'device_vs_public_max_abs': float(max(abs(vj - vj_dev).max(), abs(vk - vk_dev).max())),
    }
    if not is_df:
        out['quartets_computed'] = h.stats()['quartets_computed']
        out['quartets_screened'] = h.stats()['quartets_screened']
    if cpu is not None:
        out['cpu_baseline'] = cpu
    print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline_df(mol, dm, c_occ, workload):
    """Reference DF J/K algebra (df_jk.get_jk, pyscf/df/df_jk.py:362-380: dsymm-like half transform + dgemm) in numpy/OpenBLAS
    on a bounded sample of auxiliary rows of a random surrogate tensor of the right shape (timing only), scaled to naux."""
    from pyscf_b200.gto.mole import make_auxmol
    ncores = os.cpu_count() or 1
    aux = make_auxmol(mol)
    naux, nao = aux.nao, mol.nao
    nocc = c_occ.shape[1]
    rows = max(8, min(naux, int(2e9 / (nao * nao * 8))))   # <= 2 GB sample
    rng = np.random.RandomState(0)
    eri1 = rng.standard_normal((rows, nao, nao))
    orbo = np.asfortranarray(c_occ * np.sqrt(2.0))
    dmtril = rng.standard_normal(nao * (nao + 1) // 2)
    packed = rng.standard_normal((rows, nao * (nao + 1) // 2))
    t = time.perf_counter()
    vj = dmtril.dot(packed.T).dot(packed)
    buf = np.matmul(eri1, orbo)   # (P, nao, nocc)
    buf = buf.transpose(0, 2, 1).reshape(-1, nao)
    vk = buf.T.dot(buf)
    dt = (time.perf_counter() - t) * naux / rows
    model = ''
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                model = line.split(':', 1)[1].strip()
                break
    except Exception:
        pass
    return {'value': dt, 'unit': 's', 'cores': ncores, 'kind': 'port',
            'sample': '%d of %d auxiliary rows of %s (reference algebra df_jk.py:362-380 on a random tensor of the same shape: '
                      'two GEMV for J, batched matmul + GEMM for K, numpy/OpenBLAS all cores), time scaled by naux/rows'
                      % (rows, naux, workload), 'cpu_model': model}


def cpu_baseline(mol, dm, workload, keep=True):
    """4-center CPU arm: the reference's own CVHFnr_direct_drv + nrs8 digestion + CVHFnrs8_prescreen compiled from the
    reference sources (oracle/_ref, kind "reference") when present, else the oracle's restatement (kind "port").
    Either way the integral function is oracle_cint.c's int2e_sph — libcint is not vendored in the reference tree."""
    from oracle import oracle as O
    from oracle import ref_driver as R
    ncores = os.cpu_count() or 1
    os.environ.setdefault('OMP_NUM_THREADS', str(ncores))
    t = time.perf_counter()
    if R.available():
        vj, vk = R.get_jk(mol, dm, hermi=1)
        kind, nq = 'reference', None
        what = ('reference driver/screening/digestion (pyscf/lib/vhf/nr_direct.c, nr_direct_dot.c, optimizer.c compiled in '
                'place) + oracle McMurchie-Davidson int2e_sph (libcint absent)')
    else:
        vj, vk, nq = O.get_jk(mol, dm, return_count=True)
        kind = 'port'
        what = 'oracle McMurchie-Davidson integrals + s8 digestion restatement, OpenMP over shell pairs'
    dt = time.perf_counter() - t
    model = ''
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                model = line.split(':', 1)[1].strip()
                break
    except Exception:
        pass
    out = {'value': dt, 'unit': 's', 'cores': ncores, 'kind': kind,
           'sample': 'one full J/K build of %s (every screened shell quartet), %s' % (workload, what), 'cpu_model': model}
    if keep:
        out['vj'], out['vk'] = vj, vk
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    mol = build_mol(w)
    dm = scf_like_dm(mol.nao, w['nocc'])
    times = []
    base = None
    for k in range(args.warmup + args.steps):
        if w['kind'] == 'df':
            rng = np.random.RandomState(1)
            c_occ, _ = np.linalg.qr(rng.standard_normal((mol.nao, w['nocc'])))
            base = cpu_baseline_df(mol, dm, c_occ, args.workload)
        else:
            base = cpu_baseline(mol, dm, args.workload, keep=False)
        if k >= args.warmup:
            times.append(base['value'])
    v = float(np.mean(times))
    base['value'] = v
    out = {'impl': 'reference', 'metric': 'J/K Fock-build wall-s/iter', 'value': v, 'unit': 's', 'n_gpus': world,
           'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': v * 1e3, 'higher_is_better': False,
           'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
           'config': {'workload': args.workload, 'molecule': w['geom'], 'basis': w['basis'], 'nao': mol.nao,
                      'path': ('DF J/K (reference algebra on the host)' if w['kind'] == 'df'
                               else '4-center direct J/K (hermi=1, with_j, with_k)'), 'direct_scf_tol': 1e-13,
                      'dm': 'SCF-like 2*C_occ*C_occ^T, orthonormal random C_occ, seed 1',
                      'parallelism': 'host cores of rank 0 (OpenMP), no GPU',
                      'note': 'reference CPU path: the reference driver/screening/digestion compiled from its own sources '
                              '(oracle/_ref) around the oracle integral function; libcint itself is not vendored in the '
                              'reference tree (DESIGN.md section 2)'},
           'cpu_baseline': base,
           'e2e': {'value': v, 'unit': 's', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='benzene-ccpvtz-direct', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        if args.steps > 3:
            args.steps = 3
        args.warmup = min(args.warmup, 1)
        run_reference(args, rank, world)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args, rank, world)


if __name__ == '__main__':
    main()
