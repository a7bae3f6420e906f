#!/usr/bin/env python
"""CPU emulation of the index arithmetic of csrc/conv3_tc_kdn.cu (experimental kd-in-N convolution): a Python port of
pack_kdn_elem, the shared-memory operand descriptors (K-major, no swizzle: element (row r, k) of an operand lives at
(k // 8) * LBO + (r // 8) * SBO + (r % 8) * 16 + (k % 8) * 2 bytes from its start) and the TMEM slot mapping, run on one
4 x 16 x 8 tile in float64 and compared with torch.conv3d.  Checks the algorithm and the layouts, not the PTX.

    python tools/kdn_emulate.py
"""
import itertools

import numpy as np
import torch
import torch.nn.functional as F

TD, TH, TW = 4, 16, 8
HD, HH, HW = 6, 18, 10


def pack_kdn(w, dgrad):
    """Python port of pack_kdn_elem: returns the flat pack (float64) and its geometry."""
    cout_l, cin_l = w.shape[0], w.shape[1]
    gin, gout = (cout_l, cin_l) if dgrad else (cin_l, cout_l)
    cin8 = gin == 8
    nmp, ngroups = (5 if cin8 else 9), 4 * gout // 8
    kslices = 1 if cin8 else gin // 16
    total = kslices * nmp * 2 * ngroups * 64
    wf = w.reshape(cout_l, cin_l, 27)
    out = np.zeros(total)
    for i in range(total):
        r = i
        ch8 = r % 8; r //= 8
        r8 = r % 8; r //= 8
        ng = r % ngroups; r //= ngroups
        kc = r % 2; r //= 2
        m = r % nmp; r //= nmp
        ks = r
        kdp, go = ng // (gout // 8), (ng % (gout // 8)) * 8 + r8
        if cin8:
            gi, t = ch8, 2 * m + kc
            if t > 8:
                t = -1
        else:
            gi, t = ks * 16 + kc * 8 + ch8, m
        if kdp > 2 or t < 0 or gi >= gin:
            continue
        tap = kdp * 9 + t
        out[i] = wf[gi, go, 26 - tap] if dgrad else wf[go, gi, tap]
    return out, (gin, gout, cin8, nmp, kslices)


def run(cin, cout, dgrad, seed):
    rng = np.random.RandomState(seed)
    w = rng.randn(cout, cin, 3, 3, 3)
    pack, (gin, gout, cin8, nmp, kslices) = pack_kdn(w, dgrad)
    N = 4 * gout
    x = rng.randn(gin, TD, TH, TW)                      # one tile = the whole volume (zero padding outside)
    halo = np.zeros((HD, HH, HW, gin))
    halo[1:1 + TD, 1:1 + TH, 1:1 + TW, :] = x.transpose(1, 2, 3, 0)
    bh, bw = HH, HW
    acc = np.zeros((128, 9 * gout))                     # one accumulator buffer: 9 slots x gout columns
    for ks in range(kslices):
        # staged A: 8-channel planes [voxel][8 ch]; plane c (c = 0, 1) of this 16-channel slice
        planes = [halo[..., ks * 16 + c * 8: ks * 16 + c * 8 + 8].reshape(-1, 8) if not cin8 or c == 0 else None for c in range(2)]
        if cin8:
            planes = [halo.reshape(-1, 8), None]
        bbase = ks * nmp * N * 32 // 2                  # element offset of this slice's B'
        for q, m in itertools.product(range(HD), range(nmp)):
            # ---- A[r][k]: r = 16 h-rows x 8 w-voxels of one d-plane, k = 16 ----
            A = np.zeros((128, 16))
            a_plane = q * bh * bw                       # voxel units
            for r in range(128):
                rg, rr = r // 8, r % 8                  # SBO = one halo row, 16 B per voxel inside a core matrix
                for k in range(16):
                    kchunk, ke = k // 8, k % 8
                    if cin8:
                        t1 = 2 * m
                        t2 = 2 * m + 1 if 2 * m + 1 < 9 else 8
                        o1 = (t1 // 3) * bw + t1 % 3
                        o2 = (t2 // 3) * bw + t2 % 3
                        lbo = (o2 - o1) if 2 * m + 1 < 9 else 1
                        v = a_plane + o1 + kchunk * lbo + rg * bw + rr
                        A[r, k] = planes[0].reshape(-1)[v * 8 + ke] if v * 8 + ke < planes[0].size else 0.0
                    else:
                        kh, kw = m // 3, m % 3
                        v = a_plane + kh * bw + kw + rg * bw + rr
                        A[r, k] = planes[kchunk][v, ke]
            # ---- B'[n][k] through its descriptor: LBO = N*16 B, SBO = 128 B ----
            B = np.zeros((N, 16))
            for n_, k in itertools.product(range(N), range(16)):
                off_bytes = m * (N * 32) + (k // 8) * (N * 16) + (n_ // 8) * 128 + (n_ % 8) * 16 + (k % 8) * 2
                B[n_, k] = pack[bbase + off_bytes // 2]
            col0 = (5 - q) * gout
            acc[:, col0:col0 + N] += A @ B.T            # accumulate=1 into slots (5-q) .. (5-q)+3
    out = np.zeros((gout, TD, TH, TW))
    for p in range(TD):
        blk = acc[:, (5 - p) * gout:(5 - p) * gout + gout]          # rows = (h, w)
        out[:, p] = blk.reshape(TH, TW, gout).transpose(2, 0, 1)
    xt = torch.from_numpy(x)[None]
    wt = torch.from_numpy(w)
    ref = F.conv_transpose3d(xt, wt, None, padding=1)[0].numpy() if dgrad else F.conv3d(xt, wt, None, padding=1)[0].numpy()
    err = np.abs(out - ref).max() / np.abs(ref).max()
    print("Cin=%d Cout=%d dgrad=%d: GEMM %d -> %d, %d MMAs/plane, rel max err %.2e %s" % (
        cin, cout, dgrad, gin, gout, nmp, err, "ok" if err < 1e-12 else "MISMATCH"))
    return err < 1e-12


if __name__ == "__main__":
    ok = True
    for cin, cout, dgrad in [(8, 8, 0), (16, 8, 0), (8, 16, 0), (32, 16, 0), (8, 8, 1), (8, 16, 1), (16, 8, 1), (16, 32, 1)]:
        ok &= run(cin, cout, dgrad, seed=cin * 100 + cout * 2 + dgrad)
    raise SystemExit(0 if ok else 1)
