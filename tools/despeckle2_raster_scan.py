#!/usr/bin/env python3
"""Prototype (CPU, numpy) of an exact, row-parallel evaluation of the reference's despeckle2 kernel (oclrect.cl:348-371) under
its raster-order schedule - the plan of DESIGN.md section 8 for removing the Jacobi canonicalisation (Q3).

In raster order a small-region pixel p takes the FIRST arg-max by region size over the ordered candidates
[new(UL), new(U), new(UR), new(L), old(p), old(R), old(DL), old(D), old(DR)].  Once the row above is final only new(L) is
unknown, and p acts on it as  f_p(X) = X if size(X) >= T_p else C_p,  with C_p the first arg-max of the known candidates and
T_p = size(C_p) + 1 if C_p comes from the candidates in front of L, size(C_p) otherwise.  Two such maps compose to one of the
two:  g o f = f if size(C_f) >= T_g else g  - so a row is a segmented scan over its runs of small-region pixels.
The script checks the row-scan result against the reference's own kernel run sequentially (oracle/_ref/librd_ref.so).
usage: despeckle2_raster_scan.py [iw ih seed...]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import oracle_lib as ol  # noqa: E402
import ref_lib as rl  # noqa: E402


def row_scan_despeckle2(label, size, thre, iw, ih):
    old = label.reshape(ih, iw)
    new = old.copy()
    small = size[old] <= thre
    BIG = np.iinfo(np.int64).max
    for y in range(ih):
        xs = np.flatnonzero(small[y])
        if xs.size == 0:
            continue
        # known candidates in scan order; position 3 (L) is the unknown unless L is not a small pixel / does not exist
        cand = np.full((9, xs.size), -1, np.int64)          # label or -1 (outside the image)
        k = 0
        for yy in (-1, 0, 1):
            for xx in (-1, 0, 1):
                Y, X = y + yy, xs + xx
                ok = (X >= 0) & (X < iw) & (0 <= Y < ih)
                src = new if yy < 0 else old                  # the row above is final; this row and the next are old
                v = np.where(ok, src[min(max(Y, 0), ih - 1), np.clip(X, 0, iw - 1)], -1)
                cand[k] = v
                k += 1
        left_unknown = (xs > 0) & small[y, np.maximum(xs - 1, 0)]
        key = np.where(cand >= 0, size[np.maximum(cand, 0)], 0).astype(np.int64)
        key_known = key.copy()
        key_known[3, left_unknown] = 0                         # the unknown does not take part in C
        # first arg-max of the known candidates; "maxSize = 0 / maxLabel = own label" start: a candidate needs size > 0
        best = key_known.argmax(axis=0)                        # argmax returns the first maximum
        ckey = key_known[best, np.arange(xs.size)]
        clab = np.where(ckey > 0, cand[best, np.arange(xs.size)], old[y, xs])
        T = np.where(best < 3, ckey + 1, ckey)                 # X (position 3) beats later candidates on ties, loses to earlier ones
        T = np.maximum(T, 1)
        # segmented scan: within a run of consecutive small pixels apply f left to right; run heads have a known L
        out = np.empty(xs.size, np.int64)
        i = 0
        while i < xs.size:
            j = i
            while j + 1 < xs.size and xs[j + 1] == xs[j] + 1:
                j += 1
            # sequential form of the scan (the composition rule "f if size(C_f) >= T_g else g" makes it a parallel scan)
            cur = clab[i]
            out[i] = cur
            for t in range(i + 1, j + 1):
                cur = cur if size[cur] >= T[t] else clab[t]
                out[t] = cur
            i = j + 1
        new[y, xs] = out
    return new.ravel().astype(np.int32)


def main():
    args = [int(a) for a in sys.argv[1:]]
    iw, ih = (args[0], args[1]) if len(args) >= 2 else (640, 480)
    seeds = args[2:] or [2, 9, 10]
    k_d2 = rl.kernel_direct("rect", "despeckle2")
    k_d2.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    L = ol.oracle()
    for seed in seeds:
        img = ol.synth_frame(iw, ih, seed)
        o = ol.OracleRect(iw, ih)
        o.gpu_task(img, img.shape[-1], 17)
        lab, size = o.buffer("buf5").copy(), o.buffer("tmp0").copy()
        o.close()
        L.ora_rect_calcSize(size.ctypes.data, lab.ctypes.data, iw, ih)
        ref = lab.copy()
        k_d2(iw, ih, ref.ctypes.data, size.ctypes.data, 16, iw, ih)
        jac = lab.copy()
        L.ora_rect_despeckle2(jac.ctypes.data, size.ctypes.data, 16, iw, ih)
        got = row_scan_despeckle2(lab, size, 16, iw, ih)
        print("%dx%d seed %d: row scan == reference raster schedule: %s   (Jacobi differs in %d pixels)"
              % (iw, ih, seed, np.array_equal(got, ref), int((jac != ref).sum())))


if __name__ == "__main__":
    main()
