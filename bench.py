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
  roofline            : the class kernels (one template, 55 instantiations) against the measured FP64
                        FMA-pipe peak (b200jk_fp64_peak micro-benchmark; the 4-center path is FP64-bound,
                        SURVEY.md §8d) and, beside it, the HBM figure (algorithmic bytes / time).
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


def algorithmic_counts(mol, opt):
    """Frozen algorithmic work of one direct build for `mol`: unique Cartesian ERIs of the
    unscreened upper bound, FP64 flops = per class n_quartet*n_cart_eri*(nprim_avg*nroots*3 + 12), bytes = D+J+K+pairs."""
    st = opt.stats()
    n = st['n_sph']
    bytes_alg = 3 * n * n * 8 + st['n_pairs'] * 48
    return bytes_alg


def run_ours(args, rank, world):
    import torch
    import ctypes
    from pyscf_b200.jk import VHFOpt
    from pyscf_b200 import lib as _lib
    w = WORKLOADS[args.workload]
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    mol = build_mol(w)
    nao = mol.nao
    dm_h = scf_like_dm(nao, w['nocc'])
    t0 = time.time()
    opt = VHFOpt(mol, direct_scf_tol=1e-13, device=local)
    setup_s = time.time() - t0
    h = opt.handle
    stream = torch.cuda.current_stream(dev)
    h.lib.b200jk_set_stream(h._h, ctypes.c_void_p(stream.cuda_stream))

    dm_d = torch.from_numpy(dm_h).to(dev)
    vj_d = torch.empty_like(dm_d)
    vk_d = torch.empty_like(dm_d)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)

    def step_device():
        rc = h.lib.b200jk_direct_jk_device(h._h, ctypes.c_void_p(dm_d.data_ptr()), 1, nao, 1,
                                           ctypes.c_void_p(vj_d.data_ptr()), ctypes.c_void_p(vk_d.data_ptr()))
        h.check(rc, 'b200jk_direct_jk_device')

    # weak scaling: every rank builds J/K for its own density (independent SCF replicas share nothing);
    # strong scaling of ONE build (quartet sharding + all-reduce) is reported by --scaling strong when built.
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
    launches = 0
    torch.cuda.synchronize()
    t_wall0 = time.time()
    for k in range(args.steps):
        flush.zero_()
        evs[k][0].record(stream)
        step_device()
        evs[k][1].record(stream)
        torch.cuda.synchronize()
        st = h.stats()
        kern_ms.append(st['ms_kernels'])
        launches += st['kernel_launches']
    torch.cuda.synchronize()
    t_wall = time.time() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    ms_per_step = float(np.mean(step_ms))
    # ---- end-to-end through the public plugin call with pinned host buffers
    dm_pin = torch.from_numpy(dm_h).pin_memory()
    dm_pin_np = dm_pin.numpy()
    for _ in range(2):
        opt.get_jk(dm_pin_np, hermi=1)
    e2e_ms = []
    for k in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t = time.perf_counter()
        vj, vk = opt.get_jk(dm_pin_np, hermi=1)
        e2e_ms.append((time.perf_counter() - t) * 1e3)
    clocks = sampler.finish()
    e2e_ms_mean = float(np.mean(e2e_ms))

    if world > 1:
        tt = torch.tensor([ms_per_step, e2e_ms_mean], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_per_step, e2e_ms_mean = float(tt[0]), float(tt[1])

    if rank != 0:
        return
    # ---- roofline of the class kernels
    peak = ctypes.c_double(0)
    h.lib.b200jk_set_stream(h._h, None)
    h.lib.b200jk_fp64_peak(h._h, ctypes.byref(peak))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    kernel_ms = float(np.mean(kern_ms))
    from pyscf_b200.flops import direct_jk_flops
    flops, n_eri = direct_jk_flops(mol)
    bytes_alg = algorithmic_counts(mol, opt)
    fp64_ach = flops / (kernel_ms * 1e-3) / 1e12 if flops else None
    roof = {'bound': 'fp64', 'achieved': fp64_ach, 'peak': peak.value, 'unit': 'TFLOP/s',
            'frac': (fp64_ach / peak.value) if (fp64_ach and peak.value) else None, 'traffic': None,
            'kernel': 'jk_class_kernel<QClass<LI,LJ,LK,LL,NP>,NQ> (all class launches of one build)',
            'kernel_ms_per_step': kernel_ms, 'alg_flops_per_step': flops, 'alg_cart_eris_per_step': n_eri,
            'peak_source': 'b200jk_fp64_peak DFMA micro-benchmark (MEASURED_PEAKS.json has no fp64 entry)',
            'hbm': {'bound': 'hbm', 'achieved': bytes_alg / (kernel_ms * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                    'frac': bytes_alg / (kernel_ms * 1e-3) / 1e9 / hbm_peak, 'alg_bytes_per_step': bytes_alg,
                    'peak_source': 'MEASURED_PEAKS.json hbm_gbs (of measured)' if peaks else 'fallback 6650'}}
    # ---- CPU baseline (oracle port), rank 0, N=1 only
    cpu = None
    if world == 1 and not args.no_cpu:
        cpu = cpu_baseline(mol, dm_h, args.workload)
        if cpu.get('vj') is not None:
            err_j = float(abs(vj - cpu.pop('vj')).max())
            err_k = float(abs(vk - cpu.pop('vk')).max())
            cpu['max_abs_dJ_vs_gpu'] = err_j
            cpu['max_abs_dK_vs_gpu'] = err_k
    out = {
        'metric': 'J/K Fock-build wall-s/iter', 'value': ms_per_step * 1e-3 / 1.0, 'unit': 's',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'higher_is_better': False, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': args.workload, 'molecule': w['geom'], 'basis': w['basis'], 'nao': nao,
                   'path': '4-center direct J/K (hermi=1, with_j, with_k)', 'direct_scf_tol': 1e-13,
                   'dm': 'SCF-like 2*C_occ*C_occ^T, orthonormal random C_occ, seed 1', 'builds_per_step': world,
                   'l2_flush': '256 MiB memset between steps, outside the per-step CUDA-event pairs',
                   'parallelism': 'replicas' if world > 1 else 'single'},
        'e2e': {'value': e2e_ms_mean * 1e-3, 'unit': 's', 'h2d_bytes_per_step': int(nao * nao * 8),
                'd2h_bytes_per_step': int(2 * nao * nao * 8), 'api': 'pyscf_b200.jk.VHFOpt.get_jk (pinned host dm)'},
        'gpu_launches': int(launches), 'setup_s': setup_s, 'clocks': clocks, 'roofline': roof,
        'wall_s_timed_region': t_wall, 'quartets_computed': h.stats()['quartets_computed'],
        'quartets_screened': h.stats()['quartets_screened'],
    }
    if cpu is not None:
        out['cpu_baseline'] = cpu
    print(json.dumps(out))


def cpu_baseline(mol, dm, workload, keep=True):
    from oracle import oracle as O
    ncores = os.cpu_count() or 1
    os.environ.setdefault('OMP_NUM_THREADS', str(ncores))
    t = time.perf_counter()
    vj, vk, nq = O.get_jk(mol, dm, return_count=True)
    dt = time.perf_counter() - t
    model = ''
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                model = line.split(':', 1)[1].strip()
                break
    except Exception:
        pass
    out = {'value': dt, 'unit': 's', 'cores': ncores, 'kind': 'port',
           'sample': 'one full J/K build of %s (all %d screened shell quartets), oracle McMurchie-Davidson '
                     'integrals + s8 digestion, OpenMP over shell pairs' % (workload, nq),
           'cpu_model': model}
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
        base = cpu_baseline(mol, dm, args.workload, keep=False)
        if k >= args.warmup:
            times.append(base['value'])
    v = float(np.mean(times))
    base['value'] = v
    out = {'impl': 'reference', 'metric': 'J/K Fock-build wall-s/iter', 'value': v, 'unit': 's', 'n_gpus': world,
           'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': v * 1e3, 'higher_is_better': False,
           'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
           'config': {'workload': args.workload, 'molecule': w['geom'], 'basis': w['basis'], 'nao': mol.nao,
                      'path': '4-center direct J/K (hermi=1)', 'direct_scf_tol': 1e-13,
                      'note': 'reference CPU path = oracle port (libcint is not vendored in the reference tree; '
                              'see DESIGN.md)'},
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
