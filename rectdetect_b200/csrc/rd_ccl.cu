// rd_ccl.cu - connected-component labelling, the primitive behind label8x (oclimgutil.cl:495-538, used three times
// per frame), labelpl (oclpolyline.cl:312-355) and labelxPreprocess + labelMergeMain (oclrect.cl:289-334).
//
// The reference iterates a label-equivalence kernel a bounded number of times (10 / 11 / 8 launches); its fixed
// point gives every pixel the smallest linear index of its component.  Here that fixed point is computed exactly
// (DESIGN.md "Canonical semantics") by a two-level union-find in which the smaller index is always
// the root:
//   1. k_ccl_tile   : one CTA per 32x32 tile; the per-pixel link mask (which of W, NW, N, NE are connected) is
//                     evaluated once and kept in a byte plane; components inside the tile are resolved entirely
//                     in shared memory (init to the smallest connected backward neighbour, compress, atomicMin
//                     unions, compress) and written out as the global index of the tile-local root;
//   2. k_ccl_seams  : pixels on tile seams unite with their connected neighbours in the adjacent tile (global
//                     atomicMin union-find; ~6 % of the pixels take part);
//   3. k_ccl_flatten: every pixel reads its final root.
// HBM traffic: predicate inputs once, 1 B/px links write + seam re-reads, label plane written twice and read once.
#include "rd_common.cuh"

#define TW 32
#define TH 32
#define CCL_THREADS 256
#define L_W 1
#define L_NW 2
#define L_N 4
#define L_NE 8
#define L_BG 0x80

// ---- link functors: which backward neighbours (W, NW, N, NE) of (x, y) belong to the same component ----
struct Link8x {                       // label8xMain_int_int: equal value, value != bgc (oclimgutil.cl:511-538)
  const int *pix; int bgc, iw, ih;
  __device__ __forceinline__ void shift(size_t o) { rd_batch_off(o, pix); }
  __device__ __forceinline__ unsigned operator()(int x, int y) const {
    const int p = y * iw + x, v = pix[p];
    if (v == bgc) return L_BG;
    unsigned m = 0;
    if (x > 0 && pix[p - 1] == v) m |= L_W;
    if (y > 0) {
      if (pix[p - iw] == v) m |= L_N;
      if (x > 0 && pix[p - iw - 1] == v) m |= L_NW;
      if (x < iw - 1 && pix[p - iw + 1] == v) m |= L_NE;
    }
    return m;
  }
};
struct LinkPl {                       // labelpl_main: numbers (+1) both non-zero and differing by at most 1 (oclpolyline.cl:325-355)
  const int *num; int iw, ih;
  __device__ __forceinline__ void shift(size_t o) { rd_batch_off(o, num); }
  __device__ __forceinline__ static bool con(int a, int b) { return b != 0 && abs(a - b) <= 1; }
  __device__ __forceinline__ unsigned operator()(int x, int y) const {
    const int p = y * iw + x, v = num[p];
    if (v == 0) return L_BG;
    unsigned m = 0;
    if (x > 0 && con(v, num[p - 1])) m |= L_W;
    if (y > 0) {
      if (con(v, num[p - iw])) m |= L_N;
      if (x > 0 && con(v, num[p - iw - 1])) m |= L_NW;
      if (x < iw - 1 && con(v, num[p - iw + 1])) m |= L_NE;
    }
    return m;
  }
};
struct LinkMerge {                    // labelxPreprocess + labelMergeMain, canonical symmetric form (see rd_rect.cu / DESIGN.md)
  const uint32_t *pix; const int *mask; const int *edge; int iw, ih;
  __device__ __forceinline__ void shift(size_t o) { rd_batch_off(o, pix, mask, edge); }
  __device__ __forceinline__ bool interior(int x, int y) const { return x > 0 && y > 0 && x < iw - 1 && y < ih - 1; }
  __device__ __forceinline__ unsigned operator()(int x, int y) const {
    const int b = y * iw + x;
    const uint32_t v = pix[b];
    unsigned m = 0;
    const bool upSame = y > 0 && pix[b - iw] == v;
    const bool e = edge[b] <= 0;
    const bool mb = mask[b] != 0;
    if (y > 0) {
      const int a = b - iw;
      if (upSame) m |= L_N;                                                     // preprocess link
      else if ((interior(x, y) || interior(x, y - 1)) && e && (mb || mask[a] != 0)) m |= L_N;
    }
    if (x > 0) {
      const int a = b - 1;
      const bool same = pix[a] == v;
      if (same && !upSame) m |= L_W;                                            // preprocess link (left only when up differs)
      else if ((interior(x, y) || interior(x - 1, y)) && e && (same || mb || mask[a] != 0)) m |= L_W;
    }
    return m;
  }
};

__device__ __forceinline__ int sm_find(volatile int *L, int x) {
  int p = L[x];
  while (p != x) { x = p; p = L[x]; }
  return x;
}
__device__ __forceinline__ void sm_unite(int *L, int a, int b) {
  for (;;) {
    a = sm_find(L, a);
    b = sm_find(L, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }
    int old = atomicMin(L + a, b);
    if (old == a) return;
    a = old;
  }
}

// local neighbour offsets for the four link bits (tile-local index = ly * TW + lx)
template <class LinkFn>
__global__ void __launch_bounds__(CCL_THREADS) k_ccl_tile(int *label, uint8_t *links, LinkFn f, int iw, int ih, size_t fs) {
  rd_batch_z(fs, label, links);
  f.shift((size_t)blockIdx.z * fs);
  __shared__ int L[TW * TH];
  __shared__ uint8_t M[TW * TH];
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int lx = threadIdx.x & 31, wy = threadIdx.x >> 5;        // 8 warps, each warp owns rows wy, wy+8, wy+16, wy+24
  const int x = x0 + lx;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int ly = wy + k * 8, y = y0 + ly, i = ly * TW + lx;
    unsigned m = L_BG, full = L_BG;
    if (x < iw && y < ih) {
      full = f(x, y);
      links[(size_t)y * iw + x] = (uint8_t)full;
      m = full;
      if (lx == 0) m &= ~(L_W | L_NW);                             // neighbours outside the tile are the seam kernel's business
      if (lx == TW - 1) m &= ~L_NE;
      if (ly == 0) m &= ~(L_NW | L_N | L_NE);
    }
    M[i] = (uint8_t)m;
    int l = i;                                                     // smallest connected backward neighbour: NW < N < NE < W
    if (m & L_W) l = i - 1;
    if (m & L_NE) l = i - TW + 1;
    if (m & L_N) l = i - TW;
    if (m & L_NW) l = i - TW - 1;
    L[i] = l;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; k++) { const int i = (wy + k * 8) * TW + lx; L[i] = sm_find(L, i); }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int i = (wy + k * 8) * TW + lx;
    const unsigned m = M[i];
    if (m & L_W) sm_unite(L, i, i - 1);
    if (m & L_NW) sm_unite(L, i, i - TW - 1);
    if (m & L_N) sm_unite(L, i, i - TW);
    if (m & L_NE) sm_unite(L, i, i - TW + 1);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int ly = wy + k * 8, y = y0 + ly, i = ly * TW + lx;
    if (x < iw && y < ih) {
      const int r = sm_find(L, i);
      label[(size_t)y * iw + x] = (y0 + (r >> 5)) * iw + x0 + (r & 31);
    }
  }
}

// unions across tile seams.  One thread per pixel of the image; only seam pixels do work.
__global__ void k_ccl_seams(int *label, const uint8_t *links, int iw, int ih, size_t fs) {
  rd_batch_z(fs, label, links);
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  const int lx = x & (TW - 1), ly = y & (TH - 1);
  if (lx != 0 && lx != TW - 1 && ly != 0) return;
  const int p = y * iw + x;
  const unsigned m = links[p];
  if (m & L_BG) return;
  if ((m & L_W) && lx == 0) rd_uf_unite(label, p, p - 1);
  if ((m & L_NW) && (lx == 0 || ly == 0)) rd_uf_unite(label, p, p - iw - 1);
  if ((m & L_N) && ly == 0) rd_uf_unite(label, p, p - iw);
  if ((m & L_NE) && (lx == TW - 1 || ly == 0)) rd_uf_unite(label, p, p - iw + 1);
}

// final labels.  mode 0: background -> bgval, others -> root.
__global__ void k_ccl_flatten(int *label, const uint8_t *links, int bgval, int n, size_t fs) {
  rd_batch_y(fs, label, links);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (links[p] & L_BG) { label[p] = bgval; return; }
  label[p] = rd_uf_find(label, p);
}
// The flatten above rewrites label[] while other threads still walk it; every intermediate value is an ancestor of
// the pixel (parents only ever move towards the root), so concurrent walks stay correct.

// labelMerge: interior pixels get the root, image-border pixels keep their labelxPreprocess value (oclrect.cl:289-298)
__global__ void k_ccl_flatten_merge(int *out, const int *label, const uint32_t *pix, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, label, pix);
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  const int p = y * iw + x;
  if (x > 0 && y > 0 && x < iw - 1 && y < ih - 1) { out[p] = rd_uf_find(label, p); return; }
  const uint32_t v = pix[p];
  int l = p;
  if (y > 0 && pix[p - iw] == v) l = p - iw;
  else if (x > 0 && pix[p - 1] == v) l = p - 1;
  out[p] = l;
}

template <class LinkFn>
static void ccl_core(int *label, uint8_t *links, LinkFn f, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  RD_LAUNCH(k_ccl_tile<LinkFn>, rd_gz(dim3(rd_cdiv(iw, TW), rd_cdiv(ih, TH)), nb), CCL_THREADS, 0, s, label, links, f, iw, ih, fs);
  const dim3 b(32, 8);
  RD_LAUNCH(k_ccl_seams, rd_gz(rd_grid2d(iw, ih, b), nb), b, 0, s, label, links, iw, ih, fs);
}

// scratch: iw*ih bytes
void rd_label8x(int *label, const int *pix, void *scratch, int bgc, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  Link8x f = {pix, bgc, iw, ih};
  ccl_core(label, (uint8_t *)scratch, f, iw, ih, nb, fs, s);
  RD_LAUNCH(k_ccl_flatten, rd_gy(rd_cdiv(iw * ih, 256), nb), 256, 0, s, label, (const uint8_t *)scratch, -1, iw * ih, fs);
}
// labelpl (oclpolyline.c:170-184): `num` already holds number+1 (0 where the number is 0); zero pixels get label 0
void rd_labelpl(int *label, const int *num, void *scratch, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  LinkPl f = {num, iw, ih};
  ccl_core(label, (uint8_t *)scratch, f, iw, ih, nb, fs, s);
  RD_LAUNCH(k_ccl_flatten, rd_gy(rd_cdiv(iw * ih, 256), nb), 256, 0, s, label, (const uint8_t *)scratch, 0, iw * ih, fs);
}
// labelxPreprocess + 8 x labelMergeMain (oclrect.c:325-331), converged.  work: iw*ih ints, scratch: iw*ih bytes; out may not alias work
void rd_labelMerge(int *out, int *work, const uint32_t *pix, const int *mask, const int *edge, void *scratch, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  LinkMerge f = {pix, mask, edge, iw, ih};
  ccl_core(work, (uint8_t *)scratch, f, iw, ih, nb, fs, s);
  const dim3 b(32, 8);
  RD_LAUNCH(k_ccl_flatten_merge, rd_gz(rd_grid2d(iw, ih, b), nb), b, 0, s, out, work, pix, iw, ih, fs);
}
