// rd_despeckle2.cuh - the exact (raster-order) despeckle2 of the reference (oclrect.cl:348-371), building blocks shared by the
// CUDA kernels (rd_despeckle2.cu) and their host replay (tests/emu_despeckle2x.cpp).
//
// The reference kernel updates the region labels IN PLACE; with its work-items in raster order (the schedule of
// oracle/_ref/librd_ref.so, and the canonical one here) a small-region pixel p = (x, y) takes the FIRST arg-max by region size
// over the ordered candidates
//     code 1..9 :  new(NW) new(N) new(NE) new(W) old(p) old(E) old(SW) old(S) old(SE)          (code 0: the start value, own label / size 0)
// Pixels of large regions (size > thre) never change, so only the *small* causal neighbours carry a dependency:
//   * "static" part : every candidate that is not a small causal neighbour is known up front -> best static candidate
//                     (label, size, code) + a 4-bit mask `dyn` of the causal neighbours that are small (bit 0 NW, 1 N, 2 NE, 3 W);
//   * row above     : once row y-1 is final the NW / N / NE candidates are known -> merged by (size, code) order;
//   * same row      : only new(W) = X is unknown, and p acts on it as  f(X) = X if size(X) >= T else C,  C = best known candidate,
//                     T = size(C) + 1 if C sits in front of W in the scan (code < 4), size(C) otherwise (W beats later candidates on
//                     ties, loses to earlier ones).  Two such maps compose to one of the two:  g o f = f if size(C_f) >= T_g else g
//                     (size(C_f) is T_f or T_f - 1, so the third zone of the general composition is empty): a run of small pixels is
//                     a scan.  A pixel without a small W neighbour is a constant ("head", T = D2_HEAD).
#ifndef RD_DESPECKLE2_CUH
#define RD_DESPECKLE2_CUH

#ifdef __CUDACC__
#define D2_HD __host__ __device__ __forceinline__
#else
#define D2_HD static inline
#endif

#define D2_HEAD 0x7fffffff
#define D2_CHUNK 32

// (l, s) as candidate `c` against the best so far: larger size wins, at equal size the earlier position of the scan
D2_HD void d2_take(int &bl, int &bs, int &code, int l, int s, int c) {
  if (s > bs || (s == bs && c < code)) { bl = l; bs = s; code = c; }
}

// static part of pixel (x, y) of a small region.  Returns the list record  x | code << 16 | dyn << 20.
D2_HD int d2_static(int x, int y, const int *label, const int *size, int thre, int iw, int ih, int &bl, int &bs) {
  const int p0 = y * iw + x;
  bl = label[p0]; bs = 0;
  int code = 0, dyn = 0, k = 0;
  for (int yy = -1; yy <= 1; yy++)
    for (int xx = -1; xx <= 1; xx++, k++) {
      if (x + xx < 0 || x + xx >= iw || y + yy < 0 || y + yy >= ih) continue;
      const int l1 = label[p0 + yy * iw + xx], s1 = size[l1];
      if (k < 4 && !(s1 > thre)) { dyn |= 1 << k; continue; }
      if (s1 > bs) { bs = s1; bl = l1; code = k + 1; }
    }
  return x | (code << 16) | (dyn << 20);
}

// threshold of the map a pixel applies to new(W)
D2_HD int d2_threshold(int bs, int code) {
  const int t = bs + (code < 4 ? 1 : 0);
  return t < 1 ? 1 : t;
}

// one step of the scan: (lL, lS, lT) is the composed map of the range to the left, (bl, bs, T) of this range (T != D2_HEAD)
D2_HD void d2_compose(int &bl, int &bs, int &T, int lL, int lS, int lT) {
  if (lS >= T) { bl = lL; bs = lS; T = lT; }
  else if (lT == D2_HEAD) T = D2_HEAD;
}

#endif
