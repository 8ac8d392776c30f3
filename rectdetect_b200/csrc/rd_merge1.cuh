// rd_merge1.cuh - the FIRST pass of labelMergeMain (oclrect.cl:300-334, after labelxPreprocess :289-298) replayed exactly in the
// raster order of the reference run (oracle/_ref), as a row wavefront.  Shared by the CUDA kernels (rd_ccl.cu) and their host replay
// (tests/emu_merge1.cpp).
//
// Why the first pass and only it: seeded with the label plane the reference holds after its first pass, the schedule-independent
// fixed point of the adopt rule (the gating rounds of rd_ccl.cu) reproduces the reference's final labels (tools/ref_vs_oracle_sweep.py:
// all 249 rectangles of the sweeps, region map identical on 31 of 33 frames); everything that depends on the order of the work-items
// happens while `og` is still the preprocess pointer - which single pixels  atomic_min(&label[og], g)  moves from one tree to another.
//
// What makes that pass replayable in parallel: with L0 = the preprocess pointer (up neighbour if it has the pixel's colour, else left
// neighbour if it has, else the pixel itself) a pixel p = (x, y) processed in raster order
//   * finds label[p] = L0[p] =: og (writes only ever go to og(w) <= w and w of pixels w in front of p), label[p+1] = L0[p+1] and
//     label[p+iw] = L0[p+iw] >= p (the lower neighbour is never adopted; the right one only as L0[p+1] = p+1-iw);
//   * writes label[og] and label[p]: og is the pixel above, the pixel to the left or p itself.  So label[q] only ever receives the
//     result of q itself, of q+1 (if L0[q+1] = q) and of q+iw (if L0[q+iw] = q) and is final one row later;
//   * chases pointers into an up-right cone: in row y-k it reads columns <= x + 2k - 1, whose labels are complete once row y-k
//     is done up to column x + 2k + 1.
// Rows can therefore run concurrently, each M1_SKEW = 4 pixels behind the row above (lane = row, the lanes of a warp in lock
// step), if the contribution of the pixel BELOW q is kept apart from the rest of label[q]:
//     A[q] = min(L0[q], result of q, result of q+1 if it points at q)        B[q] = result of q+iw if it points at q (else M1_NONE)
// A reader at p sees label[q] = q < p - iw ? min(A[q], B[q]) : A[q]  (the pixel below q comes before p in raster order iff q + iw < p) -
// rows further down that are already under way write B of rows the reader must not see yet, and nothing else.
// After the pass label[q] = min(A[q], B[q]).
#ifndef RD_MERGE1_CUH
#define RD_MERGE1_CUH

#ifdef __CUDACC__
#define M1_HD __host__ __device__ __forceinline__
#define M1_MEMBER __host__ __device__ __forceinline__
#else
#define M1_HD static inline
#define M1_MEMBER inline
#endif

#define M1_SKEW 4
#define M1_NONE 0x7fffffff
// the per-pixel record of k_m1_pre (one byte)
#define M1_U 1u                         // may adopt from the pixel above  ((same colour || mask) && edge[p] <= 0)
#define M1_L 2u                         // ... from the left
#define M1_R 4u                         // ... from the right, whose label is still L0[p+1] = p+1-iw  (same colour || mask) && edge[p+1] <= 0 && pix[p+1] == pix[p+1-iw]
#define M1_KIND_SHIFT 3                 // og: 0 = p itself, 1 = the pixel above, 2 = the pixel to the left
#define M1_INT 32u                      // not on the image frame (the kernel skips frame pixels)

// the record of pixel (x, y); c = pix[p], cu / cl / cr / cru = colours of the upper / left / right / upper right neighbour (used only
// where they exist), m = mask[p] != 0, e0 = edge[p] <= 0, e1 = edge[p+1] <= 0
M1_HD unsigned m1_record(int x, int y, int iw, int ih, unsigned c, unsigned cu, unsigned cl, unsigned cr, unsigned cru, bool m, bool e0, bool e1, int &L0) {
  const int p = y * iw + x;
  const bool su = y > 0 && cu == c, sl = x > 0 && cl == c;
  const unsigned kind = su ? 1u : (sl ? 2u : 0u);
  L0 = su ? p - iw : (sl ? p - 1 : p);
  if (!(x > 0 && y > 0 && x < iw - 1 && y < ih - 1)) return kind << M1_KIND_SHIFT;
  unsigned f = M1_INT | (kind << M1_KIND_SHIFT);
  if ((su || m) && e0) f |= M1_U;
  if ((sl || m) && e0) f |= M1_L;
  if ((cr == c || m) && e1 && cr == cru) f |= M1_R;
  return f;
}

// ---- the time-major layout of the planes while the wavefront runs.  Lane k of a group of R <= 32 rows (rows 32 G + k) does pixel
// x at step t = x + M1_SKEW * k, so element (x, y) lives at  32 G iw + ((x + M1_SKEW k) mod iw) R + k :  what the lanes of a warp
// store in one step is one contiguous run of R elements (lane = row in the image layout would be 32 scattered sectors per access, and
// freshly written lines are not in L1).  A bijection of the frame's iw * ih elements: no padding.
M1_HD int m1_wrap(int v, int iw) {
  if (v >= iw) { v -= iw; if (v >= iw) v %= iw; }
  return v;
}
M1_HD int m1_rows(int G, int ih) { const int r = ih - (G << 5); return r < 32 ? r : 32; }
M1_HD int m1_index_xy(int x, int y, int iw, int ih) {
  const int G = y >> 5, k = y & 31;
  return (G << 5) * iw + m1_wrap(x + M1_SKEW * k, iw) * m1_rows(G, ih) + k;
}
M1_HD int m1_index(int q, int iw, int ih) {
  const int y = q / iw;
  return m1_index_xy(q - y * iw, y, iw, ih);
}
// the same for frames of at least M1_BIG columns and fewer than 2^24 pixels (every video size): the row by a multiply-high,
// the wrap by one conditional subtraction
#define M1_BIG 256
// q / iw for 0 <= q < 2^24, iw >= 256: (q * M) >> (32 + L) with L = floor(log2 iw), M = ceil(2^(32+L) / iw) (< 2^32 as iw > 2^L or M = 2^32 - ... :
// for a power of two iw the host passes L - 1).  The error term q * (M * iw - 2^(32+L)) / (iw * 2^(32+L)) < q / 2^(32+L) < 1 / iw.
struct M1Div { unsigned M; int L; };
M1_HD M1Div m1_div_make(int iw) {
  int L = 0;
  while ((2 << L) < iw) L++;                                        // 2^L < iw <= 2^(L+1)
  M1Div d;
  d.L = L;
  d.M = (unsigned)((((unsigned long long)1 << (32 + L)) + (unsigned)iw - 1) / (unsigned)iw);
  return d;
}
M1_HD int m1_div(int q, M1Div d) {
#ifdef __CUDA_ARCH__
  return (int)(__umulhi((unsigned)q, d.M) >> d.L);
#else
  return (int)((((unsigned long long)(unsigned)q * d.M) >> 32) >> d.L);
#endif
}
M1_HD int m1_index_big(int q, int iw, int ih, M1Div dv) {
  const int y = m1_div(q, dv);
  const int x = q - y * iw;
  const int G = y >> 5, k = y & 31;
  int tm = x + M1_SKEW * k;
  if (tm >= iw) tm -= iw;
  return (G << 5) * iw + tm * m1_rows(G, ih) + k;
}

// memory policies of m1_pixel: how the pixel's own A, the A of its left neighbour, the B of its upper neighbour and an arbitrary
// node of the chase are reached.  M1Linear: planes in image layout (the host replay in its plain form).  M1TimeMajor: the kernel.
struct M1Linear {
  int *A, *B; int iw, p;
  M1_MEMBER int look(int q, bool withB) const { int v = A[q]; if (withB) { const int b = B[q]; if (b < v) v = b; } return v; }
  M1_MEMBER void setSelf(int v) { A[p] = v; }
  M1_MEMBER void setLeft(int v) { A[p - 1] = v; }
  M1_MEMBER void setUp(int v) { B[p - iw] = v; }
};
// The kernel's policy: positions are derived from the step (tm = t mod iw, the same for every lane) only where they are used.
// gbase = 32 G iw + k; the row above lives at ubase + ((tm + ushift) mod iw) * uR (the lane above, or lane 31 of the group above).
template <bool BIG>
struct M1TimeMajor {
  int *A, *B; int iw, ih; M1Div dv;
  int gbase, R, ubase, uR, ushift, tm;
  M1_MEMBER int wrap(int v) const { if (BIG) return v >= iw ? v - iw : v; return m1_wrap(v, iw); }
  M1_MEMBER int look(int q, bool withB) const {
    const int i = BIG ? m1_index_big(q, iw, ih, dv) : m1_index(q, iw, ih);
    int v = A[i];
    if (withB) { const int b = B[i]; if (b < v) v = b; }
    return v;
  }
  M1_MEMBER int self() const { return gbase + tm * R; }
  M1_MEMBER void setSelf(int v) { A[gbase + tm * R] = v; }
  M1_MEMBER void setLeft(int v) { A[gbase + (tm > 0 ? tm - 1 : iw - 1) * R] = v; }
  M1_MEMBER void setUp(int v) { B[ubase + wrap(tm + ushift) * uR] = v; }
};
// the fields of M1TimeMajor for row y (host replay and kernel set-up)
template <bool BIG>
M1_HD void m1_row_setup(M1TimeMajor<BIG> &m, int y) {
  const int G = y >> 5, k = y & 31;
  m.R = m1_rows(G, m.ih);
  m.gbase = (G << 5) * m.iw + k;
  if (k > 0) { m.ubase = m.gbase - 1; m.uR = m.R; m.ushift = m.wrap(m.iw - m.wrap(M1_SKEW)); }
  else { m.ubase = ((G - 1) << 5) * m.iw + 31; m.uR = 32; m.ushift = m.wrap(M1_SKEW * 31); }
}

// A label value travels between lanes with a flag in bit 31: M1_FINAL = it was a root when it was produced AND nothing can change
// that any more.  A node g read at pixel p with g < p - iw has had all three of its writers (g, g+1, g+iw) pass, so a root found
// below that limit - with B included in the look - stays a root for every later reader: no need to look again.
#define M1_FINAL 0x80000000u

// per-row state a lane carries along its row
struct M1Row {
  int gleft;                            // label of the pixel to the left as it is now
  unsigned gleftF;                      // M1_FINAL if that label is a final root
  int croot[2];                         // final roots met last, or -1
};

// one interior pixel p (record f, `aupx` = A[p - iw] as it is now, | M1_FINAL if the lane above knows it to be a final root).
// Returns the final value of A[p - 1] (nothing but p itself could still have lowered it), | M1_FINAL if it is a final root -
// what the row below will read as its `aupx` three steps later.
template <class Mem>
M1_HD unsigned m1_pixel(int p, int iw, unsigned f, unsigned aupx, Mem &mem, M1Row &r) {
  const unsigned kind = (f >> M1_KIND_SHIFT) & 3u;
  const int og = p - ((f & (1u << M1_KIND_SHIFT)) ? iw : 0) - (int)((f >> (M1_KIND_SHIFT + 1)) & 1u);   // kind 1: p - iw, kind 2: p - 1
  const int aup = (int)(aupx & ~M1_FINAL);
  const bool aupF = (aupx & M1_FINAL) != 0;
  int g = og;
  unsigned fin = (unsigned)r.gleft | r.gleftF;
  if ((f & M1_U) && aup < g) g = aup;
  if ((f & M1_L) && r.gleft < g) g = r.gleft;
  if ((f & M1_R) && p + 1 - iw < g) g = p + 1 - iw;
  const int lim = p - iw;
  bool final = false;
  for (int j = 0; j < 8; j++) {                                     // for (j < 8) g = label[g]
    if ((g == aup && aupF) || (g == r.gleft && r.gleftF) || g == r.croot[0] || g == r.croot[1]) { final = true; break; }
    // the labels of the pixel itself, of the pixel above and of the pixel to the left are in registers (their B is not visible yet)
    const int v = g == p ? p : (g == lim ? aup : (g == p - 1 ? r.gleft : mem.look(g, g < lim)));
    if (v == g) { final = g < lim; break; }
    g = v;
  }
  if (final && g != r.croot[0] && g != r.croot[1]) { r.croot[1] = r.croot[0]; r.croot[0] = g; }
  const unsigned gx = (unsigned)g | (final ? M1_FINAL : 0u);
  if (g != og) {                                                    // atomic_min(&label[og], g); atomic_min(&label[p], g)
    if (kind == 1u) mem.setUp(g);                                   // the only writer of B[p - iw]
    else if (kind == 2u) { if (g < r.gleft) { mem.setLeft(g); fin = gx; } }   // A[p - 1] is r.gleft
    mem.setSelf(g);
  }
  r.gleft = g;
  r.gleftF = final ? M1_FINAL : 0u;
  return fin;
}
#endif
