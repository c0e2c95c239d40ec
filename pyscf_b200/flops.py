"""Algorithmic FP64 work of one 4-center direct J/K build (the figure bench.py's roofline uses).

Unit = one unique Cartesian ERI of a shell quartet (ij|kl), i>=j, k>=l, (ij)>=(kl), over SEGMENTED shells
(general contractions split, as the device does).  Per unit:
    generation : n_primitive_quartets * nroots * 3 flop   (one mul + one fma per root: gx*gy*gz accumulate)
    digestion  : 12 flop                                  (2 J + 4 K fused multiply-adds, s8 symmetry;
                                                           the reference's count)
Recurrences, Rys roots and screening are overhead, not counted (they are what the kernel should
minimise).  No screening is assumed (upper bound; benzene/cc-pVTZ loses 3 % to screening).
"""
import numpy as np


def _ncart(l):
    return (l + 1) * (l + 2) // 2


def direct_jk_flops(mol):
    ls, nps = [], []
    for b in mol._bas:
        l, nprim, nctr = int(b[1]), int(b[2]), int(b[3])
        coef = mol._env[b[6]:b[6] + nprim * nctr].reshape(nctr, nprim)
        for c in range(nctr):
            ls.append(l)
            nps.append(int(np.count_nonzero(coef[c])))
    ls, nps = np.array(ls), np.array(nps)
    lmax = ls.max()
    cls = {}
    for la in range(lmax + 1):
        for lb in range(la + 1):
            na, nb = nps[ls == la], nps[ls == lb]
            if la != lb:
                npair = len(na) * len(nb)
                s1 = int(na.sum()) * int(nb.sum())
                s2 = int((na ** 2).sum()) * int((nb ** 2).sum())
            else:
                npair = len(na) * (len(na) + 1) // 2
                s1 = (int(na.sum()) ** 2 + int((na ** 2).sum())) // 2
                s2 = (int((na ** 2).sum()) ** 2 + int((na ** 4).sum())) // 2
            cls[(la, lb)] = (npair, s1, s2)
    keys = sorted(cls, key=lambda k: k[0] * (k[0] + 1) // 2 + k[1])
    flops = 0.0
    neri = 0.0
    for ib, kb in enumerate(keys):
        for kk in keys[:ib + 1]:
            nb_, sb1, sb2 = cls[kb]
            nk_, sk1, sk2 = cls[kk]
            if nb_ == 0 or nk_ == 0:
                continue
            if kb != kk:
                nq = nb_ * nk_
                pq = sb1 * sk1
            else:
                nq = nb_ * (nb_ + 1) // 2
                pq = (sb1 * sb1 + sb2) // 2
            ncomp = _ncart(kb[0]) * _ncart(kb[1]) * _ncart(kk[0]) * _ncart(kk[1])
            nroots = (kb[0] + kb[1] + kk[0] + kk[1]) // 2 + 1
            flops += ncomp * (pq * nroots * 3.0 + nq * 12.0)
            neri += ncomp * nq
    return flops, neri
