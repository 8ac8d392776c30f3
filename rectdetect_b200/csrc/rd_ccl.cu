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
//
// The merge labelling (labelMergeMain, whose outcome depends on the order of the reference's work-items) adds to this: the gating
// rounds of its one-directional pairs (k_merge_gate / k_merge_apply / k_merge_roots) and, as a mode (rd_set_merge_replay), the exact
// replay of the reference's first pass as a row wavefront (k_m1_pre / k_m1_wave / k_m1_fold / k_merge_seed, logic in rd_merge1.cuh).
#include "rd_common.cuh"
#include "rd_merge1.cuh"

#define TW 32
#define TH 32
#define CCL_THREADS 256
#define L_W 1
#define L_NW 2
#define L_N 4
#define L_NE 8
#define L_BG 0x80
#define L_ROOT 0x40                     // the pixel is the root of its component inside its tile

// ---- link functors.  load(p) reads what the predicate needs to know about ONE pixel (p = y * iw + x) (each pixel of the tile and
// of its one-pixel apron above / beside is loaded once into shared memory); link(...) then decides, from the staged
// values of a pixel (c) and of its W, NW, N, NE neighbours, which of the four belong to the same component. ----
template <class PIX>
struct Link8x {                       // label8xMain_int_int: equal value, value != bgc (oclimgutil.cl:511-538); PIX = int or uint8_t plane
  typedef int V;
  const PIX *pix; int bgc, iw, ih;
  __device__ __forceinline__ void shift(size_t o) { rd_batch_off(o, pix); }
  __device__ __forceinline__ V load(int p) const { return (int)pix[p]; }
  __device__ __forceinline__ unsigned link(V c, V w, V nw, V n, V ne, int x, int y) const {
    if (c == bgc) return L_BG;
    unsigned m = 0;
    if (x > 0 && w == c) m |= L_W;
    if (y > 0) {
      if (n == c) m |= L_N;
      if (x > 0 && nw == c) m |= L_NW;
      if (x < iw - 1 && ne == c) m |= L_NE;
    }
    return m;
  }
};
struct LinkPl {                       // labelpl_main: numbers (+1) both non-zero and differing by at most 1 (oclpolyline.cl:325-355)
  typedef int V;
  const int *num; int iw, ih;
  __device__ __forceinline__ void shift(size_t o) { rd_batch_off(o, num); }
  __device__ __forceinline__ V load(int p) const { return num[p]; }
  __device__ __forceinline__ static bool con(int a, int b) { return b != 0 && abs(a - b) <= 1; }
  __device__ __forceinline__ unsigned link(V c, V w, V nw, V n, V ne, int x, int y) const {
    if (c == 0) return L_BG;
    unsigned m = 0;
    if (x > 0 && con(c, w)) m |= L_W;
    if (y > 0) {
      if (con(c, n)) m |= L_N;
      if (x > 0 && con(c, nw)) m |= L_NW;
      if (x < iw - 1 && con(c, ne)) m |= L_NE;
    }
    return m;
  }
};
// labelxPreprocess + labelMergeMain (DESIGN.md Q6'): the schedule-independent fixed point of the adopt rule, computed here.  By default
// it starts from the preprocess pointers (they are plain links, LinkMerge::pre); with the first-pass replay on (rd_set_merge_replay /
// RD_MERGE_REPLAY=1) the reference's first pass is replayed exactly (rd_merge1.cuh, kernels below) and the fixed point starts from
// what that pass left - the reference's own region map on all but one frame of the sweeps, at ten times the cost of the labelling.  For a 4-neighbour pair (a, b), b = a+1 or a+iw, with
// edge[b] <= 0, b may adopt from a iff b is interior and (same colour or mask[b]); a may adopt from b iff a is interior and (same
// colour or mask[a]).  Pairs that may adopt in both directions are plain links of the labelling below (whatever the order, one of the
// two labels is the smaller one), the pointers the first pass left are united with it afterwards (k_merge_seed); a pair that may
// adopt in ONE direction only is marked (L_DW / L_DN: the pixel's pair with its W / N neighbour) and decided by the gating rounds
// (k_merge_gate / k_merge_apply): united when the source's label is smaller than the adopter's.
#define L_DW 0x10                       // merge only: the pair (W neighbour, this pixel) may adopt in one direction
#define L_DN 0x20                       // ... the pair (N neighbour, this pixel)
template <class MASK>
struct LinkMerge {
  struct V { uint32_t pix; uint32_t fl; };        // fl bit 0: mask != 0, bit 1: edge <= 0
  const uint32_t *pix; const MASK *mask; const int *edge; int iw, ih;
  int pre;                                        // 1: the preprocess pointers are links of the labelling too (no first-pass replay)
  __device__ __forceinline__ void shift(size_t o) { rd_batch_off(o, pix, mask, edge); }
  __device__ __forceinline__ V load(int p) const {
    V v; v.pix = pix[p]; v.fl = (mask[p] != 0 ? 1u : 0u) | (edge[p] <= 0 ? 2u : 0u);
    return v;
  }
  __device__ __forceinline__ bool interior(int x, int y) const { return x > 0 && y > 0 && x < iw - 1 && y < ih - 1; }
  // pair (a, b), b = this pixel: 3 = both directions, 1 = only b adopts from a, 2 = only a adopts from b, 0 = none
  __device__ __forceinline__ unsigned pair(V a, V b, bool ia, bool ib) const {
    if (!(b.fl & 2)) return 0;
    const bool same = a.pix == b.pix;
    return ((ib && (same || (b.fl & 1))) ? 1u : 0u) | ((ia && (same || (a.fl & 1))) ? 2u : 0u);
  }
  __device__ __forceinline__ unsigned link(V c, V w, V nw, V n, V ne, int x, int y) const {
    unsigned m = 0;
    const bool upSame = pre && y > 0 && n.pix == c.pix;
    const bool ic = interior(x, y);
    if (y > 0) {
      const unsigned d = pair(n, c, interior(x, y - 1), ic);
      if (upSame || d == 3) m |= L_N;                                           // preprocess link / both directions
      else if (d) m |= L_DN;
    }
    if (x > 0) {
      const unsigned d = pair(w, c, interior(x - 1, y), ic);
      if ((pre && w.pix == c.pix && !upSame) || d == 3) m |= L_W;               // preprocess link (left only when up differs) / both directions
      else if (d) m |= L_DW;
    }
    return m;
  }
};

// find with path halving: every second node on the way is re-pointed at its grandparent.  A stale write can only
// replace a parent by one of its ancestors (parents only ever move towards the root), so concurrent walks stay valid.
__device__ __forceinline__ int sm_find(volatile int *L, int x) {
  int p = L[x];
  while (p != x) {
    const int g = L[p];
    if (g == p) return p;
    L[x] = g;
    x = g;
    p = L[x];
  }
  return x;
}
// read-only walk, for the phase in which every run start is being pointed at its root: a concurrent path-halving write
// could put an older ancestor back over a root that was just stored
__device__ __forceinline__ int sm_find_ro(const volatile int *L, int x) {
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

// out of line: unions are rare events in the row sweep below, which is unrolled 32 times
__device__ __noinline__ void sm_unite_rare(int *L, int a, int b) { sm_unite(L, a, b); }

// neighbour exchange inside a warp for the staged values (int, or the two words of LinkMerge::V)
__device__ __forceinline__ int ccl_up1(int v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ int ccl_down1(int v) { return __shfl_down_sync(0xffffffffu, v, 1); }
template <class V> __device__ __forceinline__ V ccl_up1(V v) { V r; r.pix = __shfl_up_sync(0xffffffffu, v.pix, 1); r.fl = __shfl_up_sync(0xffffffffu, v.fl, 1); return r; }
template <class V> __device__ __forceinline__ V ccl_down1(V v) { V r; r.pix = __shfl_down_sync(0xffffffffu, v.pix, 1); r.fl = __shfl_down_sync(0xffffffffu, v.fl, 1); return r; }

// One WARP per 32x32 tile (eight tiles per CTA), lane = column, rows top to bottom - the classic two-pass labelling with the
// first pass done a row at a time by a warp:
//  - the predicate inputs are loaded once per pixel (coalesced, one row ahead) and handed to the neighbours by shuffles;
//    the row above stays in registers, lanes 0 / 31 fetch the columns beside the tile; the link masks go to the byte plane;
//  - a horizontal run is a ballot away; every pixel takes the smallest provisional label among the pixels of the row above
//    it is linked to (its own index if there is none), and a segmented min-scan gives the run its label.  A run
//    that starts a component therefore gets its first pixel's index - the smallest index of the component so far;
//  - only where a run joins DIFFERENT labels from above (the bottom of a "U") is a union recorded, in a union-find over the
//    provisional labels kept in the tile's label array itself (A[i] = label of pixel i <= i, a forest by construction).
//    These events are rare, so the row loop is almost branch-free - the previous kernel spent most of its time in the
//    unions of every run overlap;
//  - second pass: every pixel follows A[] to its root (depth 1 unless unions happened), roots are marked L_ROOT.
// Links towards pixels outside the tile are left to the seam kernel.
#define CCL_WARPS 8
template <class LinkFn>
__global__ void __launch_bounds__(CCL_WARPS * 32) k_ccl_tile(int *label, uint8_t *links, LinkFn f, int iw, int ih, int tilesX, int tiles, size_t fs) {
  rd_batch_z(fs, label, links);
  f.shift((size_t)blockIdx.z * fs);
  typedef typename LinkFn::V V;
  __shared__ int A_all[CCL_WARPS][TW * TH];
  const int lx = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int tile = blockIdx.x * CCL_WARPS + wp;
  if (tile >= tiles) return;
  int *A = A_all[wp];
  const int x0 = (tile % tilesX) * TW, y0 = (tile / tilesX) * TH;
  const int x = x0 + lx;
  const bool okx = x < iw, okw = x0 > 0, oke = x0 + TW < iw;
  const int rows = min(TH, ih - y0);
  int p = y0 * iw + x;                                             // linear index of (x, y0)
  V up = V(), upW = V(), upE = V();
  if (y0 > 0) {
    if (okx) up = f.load(p - iw);
    upW = ccl_up1(up); upE = ccl_down1(up);
    if (lx == 0 && okw) upW = f.load(p - iw - 1);
    if (lx == 31 && oke) upE = f.load(p - iw + 1);
  }
  // row 0, loaded ahead
  V cn = V(), wn = V(), en = V();
  if (okx) cn = f.load(p);
  if (lx == 0 && okw) wn = f.load(p - 1);
  if (lx == 31 && oke) en = f.load(p + 1);
  int prevLab = 0;
  unsigned fgbits = 0;                                             // bit ly: my pixel of row ly is foreground
  for (int ly = 0; ly < rows; ly++, p += iw) {
    const int y = y0 + ly, i = ly * TW + lx;
    const V c = cn;
    V w = ccl_up1(c), e = ccl_down1(c);
    if (lx == 0) w = wn;
    if (lx == 31) e = en;
    if (ly + 1 < rows) {                                           // next row
      if (okx) cn = f.load(p + iw);
      if (lx == 0 && okw) wn = f.load(p + iw - 1);
      if (lx == 31 && oke) en = f.load(p + iw + 1);
    }
    unsigned m = L_BG;
    if (okx) {
      const unsigned full = f.link(c, w, upW, up, upE, x, y);
      links[p] = (uint8_t)full;
      m = full;
      if (lx == 0) m &= ~(L_W | L_NW);                             // neighbours outside the tile are the seam kernel's business
      if (lx == TW - 1) m &= ~L_NE;
      if (ly == 0) m &= ~(L_NW | L_N | L_NE);
    }
    if (__ballot_sync(0xffffffffu, !(m & L_BG)) == 0) { up = c; upW = w; upE = e; continue; }   // nothing but background in this row
    const unsigned starts = ~__ballot_sync(0xffffffffu, (m & L_W) != 0);      // bit j set: pixel j starts a run
    const int labW = __shfl_up_sync(0xffffffffu, prevLab, 1), labE = __shfl_down_sync(0xffffffffu, prevLab, 1);
    // smallest label among the pixels of the row above this pixel is linked to (its own index if none), then the smallest
    // of the run: a segmented min-scan towards the right and a broadcast from the run's last lane.  (redux.sync with one
    // member mask per run is serialised run by run; the scan costs the same whatever the row looks like.)
    int v = i;
    if (m & L_N) v = min(v, prevLab);
    if (m & L_NW) v = min(v, labW);
    if (m & L_NE) v = min(v, labE);
    const int s = 31 - __clz(starts & (0xffffffffu >> (31 - lx)));            // first lane of my run (bit 0 of `starts` is always set)
    const unsigned right = lx == 31 ? 0u : (starts & (0xffffffffu << (lx + 1)));
    const int last = right ? __ffs(right) - 2 : 31;                           // its last lane
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, d);
      if (lx - d >= s) v = min(v, t);
    }
    const int lab = __shfl_sync(0xffffffffu, v, last);
    if (!(m & L_BG)) {
      // the run joins everything it is linked to above: whatever carries another label is united with the run's label
      if ((m & L_N) && prevLab != lab) sm_unite_rare(A, prevLab, lab);
      if ((m & L_NW) && labW != lab) sm_unite_rare(A, labW, lab);
      if ((m & L_NE) && labE != lab) sm_unite_rare(A, labE, lab);
      A[i] = lab;
      fgbits |= 1u << ly;
    }
    __syncwarp();
    // carry the CURRENT root of the label down: after two labels have been united, the rows below would otherwise keep
    // meeting the stale pair and ask for the same union again (every diagonal step of a thin string joins the two
    // background sides through the 8-neighbourhood)
    prevLab = lab;
    if (!(m & L_BG)) { const int r = A[lab]; if (r != lab) prevLab = sm_find_ro(A, r); }
    up = c; upW = w; upE = e;
  }
  // second pass: roots
  p = y0 * iw + x;
  for (int ly = 0; ly < rows; ly++, p += iw) {
    const int i = ly * TW + lx;
    if (!((fgbits >> ly) & 1u)) continue;
    const int r = sm_find_ro(A, i);
    label[p] = (y0 + (r >> 5)) * iw + x0 + (r & 31);
    if (r == i) links[p] |= L_ROOT;
  }
}

// unions across tile seams: one CTA of 96 threads per tile - warp 0 walks the top row (links to the tile row above),
// warp 1 the left column (W / NW links to the tile on the left), warp 2 the right column (NE links to the tile on the right).
__global__ void __launch_bounds__(96) k_ccl_seams(int *label, const uint8_t *links, int iw, int ih, size_t fs) {
  rd_batch_z(fs, label, links);
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int x, y;
  if (w == 0) { x = x0 + lane; y = y0; }
  else if (w == 1) { x = x0; y = y0 + lane; }
  else { x = x0 + TW - 1; y = y0 + lane; }
  const bool in = x < iw && y < ih;
  const int p = y * iw + x;
  unsigned m = in ? links[p] : (unsigned)L_BG;
  if (m & L_BG) m = 0;
  if (w == 1) { m &= L_W | L_NW; if (lane == 0) m &= ~L_NW; }                   // lane 0 is the corner pixel, its NW link is warp 0's
  else if (w == 2) { m &= L_NE; if (lane == 0) m = 0; }
  else m &= L_NW | L_N | L_NE;
  // Along a seam most pixels repeat the union their neighbour on the seam already asks for: both sides of a long run sit in the
  // same two tile components (label[] still holds tile roots here).  So a lane first looks up the two tile roots of each of its
  // links and only goes to the union-find when the pair differs from the one the previous lane holds for the same direction -
  // on the background component of a string image that removes nine unions out of ten, and with them the atomics on its hot root.
  const int a = m ? __ldcg(label + p) : -1;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const unsigned bit = k == 0 ? L_W : (k == 1 ? L_NW : (k == 2 ? L_N : L_NE));
    const int q = p + (k == 0 ? -1 : (k == 1 ? -iw - 1 : (k == 2 ? -iw : -iw + 1)));
    const bool has = (m & bit) != 0;
    const int b = has ? __ldcg(label + q) : -1;
    const int pa = __shfl_up_sync(0xffffffffu, a, 1), pb = __shfl_up_sync(0xffffffffu, b, 1);
    if (has && !(lane > 0 && pa == a && pb == b) && a != b) rd_uf_unite(label, a, b);
  }
}

// Final labels, in two steps.  k_ccl_roots: the tile roots (marked L_ROOT; a few per tile) walk to their final root and
// remember it.  Every other pixel still points at its tile root - the seam unions only ever re-parent roots - so the
// flatten kernels below need exactly one hop: label[label[p]].  (For a tile root, label[p] is already final and the final
// root points at itself.)  Rewriting label[] in place is safe: a tile root is rewritten with the value it already holds.
__global__ void k_ccl_roots(int *label, const uint8_t *links, int n, size_t fs) {
  rd_batch_y(fs, label, links);
  const int p4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (p4 >= n) return;
  unsigned m4;
  if (p4 + 3 < n) m4 = *(const uint32_t *)(links + p4);
  else { m4 = 0; for (int k = 0; p4 + k < n; k++) m4 |= (unsigned)links[p4 + k] << (8 * k); }
  if (!(m4 & 0x40404040u)) return;
#pragma unroll
  for (int k = 0; k < 4; k++)
    if ((m4 >> (8 * k)) & L_ROOT) label[p4 + k] = rd_uf_find(label, p4 + k);
}
// background -> bgval, others -> root
__global__ void k_ccl_flatten(int *label, const uint8_t *links, int bgval, int n, size_t fs) {
  rd_batch_y(fs, label, links);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (links[p] & L_BG) { label[p] = bgval; return; }
  label[p] = __ldcg(label + __ldcg(label + p));
}

// flatten + compaction: foreground pixels are also appended to `list` (list[0] = count, entries from 1, any order)
__global__ void k_ccl_flatten_list(int *label, const uint8_t *links, int *list, int bgval, int n, size_t fs) {
  rd_batch_y(fs, label, links, list);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  bool fg = false;
  if (p < n) {
    if (links[p] & L_BG) label[p] = bgval;
    else { label[p] = __ldcg(label + __ldcg(label + p)); fg = true; }
  }
  const unsigned m = __ballot_sync(0xffffffffu, fg);
  if (m == 0) return;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(list, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (fg) list[1 + base + __popc(m & ((1u << lane) - 1))] = p;
}
// labelMerge: interior pixels get the root, image-border pixels keep their labelxPreprocess value (oclrect.cl:289-298)
__global__ void k_ccl_flatten_merge(int *out, const int *label, const uint32_t *pix, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, label, pix);
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  const int p = y * iw + x;
  if (x > 0 && y > 0 && x < iw - 1 && y < ih - 1) { out[p] = __ldcg(label + __ldcg(label + p)); return; }
  const uint32_t v = pix[p];
  int l = p;
  if (y > 0 && pix[p - iw] == v) l = p - iw;
  else if (x > 0 && pix[p - 1] == v) l = p - 1;
  out[p] = l;
}

// labelMerge: interior pixels get the root.  Image-frame pixels never run the main pass: they keep what the first pass left in them
// (first[]), except that a frame pixel that was still the root of its tree then has been hooked under a smaller root since.
__global__ void k_ccl_flatten_merge1(int *out, const int *label, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, label);
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  const int p = y * iw + x;
  const int first = out[p];
  if ((x > 0 && y > 0 && x < iw - 1 && y < ih - 1) || first == p) out[p] = __ldcg(label + __ldcg(label + p));
}

template <class LinkFn>
static void ccl_core(int *label, uint8_t *links, LinkFn f, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  const int tilesX = rd_cdiv(iw, TW), tiles = tilesX * rd_cdiv(ih, TH);
  RD_LAUNCH(k_ccl_tile<LinkFn>, dim3(rd_cdiv(tiles, CCL_WARPS), 1, nb), CCL_WARPS * 32, 0, s, label, links, f, iw, ih, tilesX, tiles, fs);
  RD_LAUNCH(k_ccl_seams, dim3(rd_cdiv(iw, TW), rd_cdiv(ih, TH), nb), 96, 0, s, label, links, iw, ih, fs);
  RD_LAUNCH(k_ccl_roots, rd_gy(rd_cdiv(rd_cdiv(iw * ih, 4), 256), nb), 256, 0, s, label, (const uint8_t *)links, iw * ih, fs);
}

// scratch: iw*ih bytes
void rd_label8x(int *label, const int *pix, void *scratch, int bgc, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  Link8x<int> f = {pix, bgc, iw, ih};
  ccl_core(label, (uint8_t *)scratch, f, iw, ih, nb, fs, s);
  RD_LAUNCH(k_ccl_flatten, rd_gy(rd_cdiv(iw * ih, 256), nb), 256, 0, s, label, (const uint8_t *)scratch, -1, iw * ih, fs);
}
// byte plane in, labels out, and the foreground pixels gathered into `list` (list[0] must be zero on entry)
void rd_label8x_u8_list(int *label, const uint8_t *pix, void *scratch, int *list, int bgc, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  Link8x<uint8_t> f = {pix, bgc, iw, ih};
  ccl_core(label, (uint8_t *)scratch, f, iw, ih, nb, fs, s);
  RD_LAUNCH(k_ccl_flatten_list, rd_gy(rd_cdiv(iw * ih, 256), nb), 256, 0, s, label, (const uint8_t *)scratch, list, -1, iw * ih, fs);
}
// same on a byte plane (the fused string clean-up kernels emit bytes)
void rd_label8x_u8(int *label, const uint8_t *pix, void *scratch, int bgc, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  Link8x<uint8_t> f = {pix, bgc, iw, ih};
  ccl_core(label, (uint8_t *)scratch, f, iw, ih, nb, fs, s);
  RD_LAUNCH(k_ccl_flatten, rd_gy(rd_cdiv(iw * ih, 256), nb), 256, 0, s, label, (const uint8_t *)scratch, -1, iw * ih, fs);
}
// labelpl (oclpolyline.c:170-184): `num` already holds number+1 (0 where the number is 0); zero pixels get label 0
void rd_labelpl(int *label, const int *num, void *scratch, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  LinkPl f = {num, iw, ih};
  ccl_core(label, (uint8_t *)scratch, f, iw, ih, nb, fs, s);
  RD_LAUNCH(k_ccl_flatten, rd_gy(rd_cdiv(iw * ih, 256), nb), 256, 0, s, label, (const uint8_t *)scratch, 0, iw * ih, fs);
}
// ---- the gating rounds of the merge labelling.  label[] holds tile roots that point at final roots (after k_ccl_roots), so the
// component label of a pixel is label[label[p]].  k_merge_gate decides every marked pair on the labels as they are (nothing is
// modified but bits of the pixel's own link byte: L_NW / L_NE are free here, the merge graph is 4-connected), k_merge_apply then
// unites.  flags[r] != 0: round r enabled something; the kernels of round r > 0 return at once when flags[r - 1] == 0.
#define RD_MERGE_ROUNDS 2            // one round applies, the second confirms: no frame of the sweeps enables anything in a second round
template <class MASK>
__global__ void k_merge_gate(const int *label, uint8_t *links, LinkMerge<MASK> f, int *flags, int round, int iw, int ih, size_t fs) {
  rd_batch_y(fs, label, links, flags);
  f.shift((size_t)blockIdx.y * fs);
  const int n = iw * ih;
  if (round > 0 && flags[round - 1] == 0) return;
  bool any = false;
  for (int p4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4; p4 < n; p4 += gridDim.x * blockDim.x * 4) {
  unsigned m4;
  if (p4 + 3 < n) m4 = *(const uint32_t *)(links + p4);
  else { m4 = 0; for (int k = 0; p4 + k < n; k++) m4 |= (unsigned)links[p4 + k] << (8 * k); }
  if (!(m4 & 0x30303030u)) continue;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const unsigned m = (m4 >> (8 * k)) & 255u;
    if (!(m & (L_DW | L_DN))) continue;
    const int p = p4 + k, x = p % iw, y = p / iw;
    const typename LinkMerge<MASK>::V c = f.load(p);
    const int rc = __ldcg(label + __ldcg(label + p));
    unsigned en = 0;
    if (m & L_DW) {
      const unsigned d = f.pair(f.load(p - 1), c, f.interior(x - 1, y), f.interior(x, y));
      const int rw = __ldcg(label + __ldcg(label + p - 1));
      if ((d == 1 && rw < rc) || (d == 2 && rc < rw)) en |= L_NW;               // (adopter = this pixel, source = W) or the other way round
    }
    if (m & L_DN) {
      const unsigned d = f.pair(f.load(p - iw), c, f.interior(x, y - 1), f.interior(x, y));
      const int rn = __ldcg(label + __ldcg(label + p - iw));
      if ((d == 1 && rn < rc) || (d == 2 && rc < rn)) en |= L_NE;
    }
    if (en) { links[p] = (uint8_t)(m | en); any = true; }
  }
  }
  if (any) flags[round] = 1;
}
__global__ void k_merge_apply(int *label, uint8_t *links, const int *flags, int round, int iw, int n, size_t fs) {
  rd_batch_y(fs, label, links, flags);
  if (flags[round] == 0) return;
  for (int p4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4; p4 < n; p4 += gridDim.x * blockDim.x * 4) {
    unsigned m4;
    if (p4 + 3 < n) m4 = *(const uint32_t *)(links + p4);
    else { m4 = 0; for (int k = 0; p4 + k < n; k++) m4 |= (unsigned)links[p4 + k] << (8 * k); }
    if (!(m4 & 0x0a0a0a0au)) continue;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const unsigned m = (m4 >> (8 * k)) & 255u;
      if (!(m & (L_NW | L_NE))) continue;
      const int p = p4 + k;
      if (m & L_NW) rd_uf_unite(label, p, p - 1);
      if (m & L_NE) rd_uf_unite(label, p, p - iw);
      links[p] = (uint8_t)(m & ~(L_NW | L_NE));
    }
  }
}
// k_ccl_roots of a gating round: only when the round united something
__global__ void k_merge_roots(int *label, const uint8_t *links, const int *flags, int round, int n, size_t fs) {
  rd_batch_y(fs, label, links, flags);
  if (flags[round] == 0) return;
  for (int p4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4; p4 < n; p4 += gridDim.x * blockDim.x * 4) {
    unsigned m4;
    if (p4 + 3 < n) m4 = *(const uint32_t *)(links + p4);
    else { m4 = 0; for (int k = 0; p4 + k < n; k++) m4 |= (unsigned)links[p4 + k] << (8 * k); }
    if (!(m4 & 0x40404040u)) continue;
#pragma unroll
    for (int k = 0; k < 4; k++)
      if ((m4 >> (8 * k)) & L_ROOT) label[p4 + k] = rd_uf_find(label, p4 + k);
  }
}
// The top row of the merge labelling (see the oracle, ora_rect.cpp labelMerge): a pixel of the top row whose lower neighbour has its
// colour and is no edge pixel ends up on the START of its run of equal colours - that is where the reference's first pass drags
// it - instead of on its left neighbour.  One warp per frame: run starts by a max-scan over chunks of 32 columns.
__global__ void __launch_bounds__(32) k_merge_toprow(int *out, const uint32_t *pix, const int *edge, int iw, int ih, size_t fs) {
  rd_batch_x(fs, out, pix, edge);
  if (ih <= 2) return;
  const int lane = threadIdx.x;
  int carry = 0;
  for (int x0 = 0; x0 < iw; x0 += 32) {
    const int x = x0 + lane;
    const uint32_t c = x < iw ? pix[x] : 0u;
    int v = (x < iw && (x == 0 || pix[x - 1] != c)) ? x : -1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v = max(v, t); }
    v = max(v, carry);
    carry = __shfl_sync(0xffffffffu, v, 31);
    if (x >= 1 && x < iw - 1 && v != x && pix[iw + x] == c && edge[iw + x] <= 0) out[x] = v;
  }
}
// ---- the first pass, exactly (rd_merge1.cuh).  k_m1_pre: per-pixel records, A = L0, B = none, all three in the time-major layout
// of the wavefront.  k_m1_wave: one CTA per frame, lane = row, warp = 32 rows, the warps take the groups of 32 rows round robin; a
// warp may run step t of its group once the group above has finished step t + 32 * M1_SKEW (its last row is then M1_SKEW pixels
// ahead of this group's first row) - progress counters in shared memory.  All planes are read and written by this one CTA, so
// plain (L1) accesses are coherent.  k_m1_fold: back to the image layout, label = min(A, B).
template <class MASK>
__global__ void k_m1_pre(int *A, int *B, uint8_t *F, LinkMerge<MASK> f, int iw, int ih, size_t fs) {
  rd_batch_z(fs, A, B, F);
  f.shift((size_t)blockIdx.z * fs);
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  const int p = y * iw + x, i = m1_index_xy(x, y, iw, ih);
  const bool in = x > 0 && y > 0 && x < iw - 1 && y < ih - 1;
  int L0;
  const unsigned r = m1_record(x, y, iw, ih, f.pix[p], y > 0 ? f.pix[p - iw] : 0u, x > 0 ? f.pix[p - 1] : 0u, in ? f.pix[p + 1] : 0u, in ? f.pix[p + 1 - iw] : 0u,
                               in && f.mask[p] != 0, in && f.edge[p] <= 0, in && f.edge[p + 1] <= 0, L0);
  A[i] = L0;
  B[i] = M1_NONE;
  F[i] = (uint8_t)r;
}
// Per step a lane needs: its record (fetched a step ahead; one contiguous run of bytes per warp), the label of the pixel above as
// the row above left it - that row's lane holds it in a register three steps after it did the pixel to the right of it, so it
// travels by shuffle through a three-deep delay line (a lane whose upper row belongs to another warp, or is the top row of the
// image, reads it from memory four steps ahead: that row is at least 128 pixels further on) - and whatever the pointer chase
// touches (gathers).  A root found below the line p - iw is final (all its writers have passed): the two met last are remembered,
// and the flag travels with the label to the lane below, so that the look is only repeated where a tree is still growing -
// 0.9 % of the pixels of a 720p frame instead of 4.9 %.  What the lanes store in a step is contiguous in the time-major layout.
#define M1_POLL 8                      // steps between two looks at the progress of the group above
template <bool BIG>
__global__ void __launch_bounds__(1024) k_m1_wave(int *A0, int *B0, const uint8_t *F0, int *err, int iw, int ih, int slack, int sleepns, size_t fs) {
  rd_batch_x(fs, A0, B0, F0, err);
  int *__restrict__ A = A0, *__restrict__ B = B0;
  const uint8_t *__restrict__ F = F0;
  __shared__ volatile int prog[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
  if (threadIdx.x < 32) prog[threadIdx.x] = 0;
  __syncthreads();
  const int groups = (ih + 31) >> 5, S = (iw + M1_SKEW * 31 + M1_POLL - 1) & ~(M1_POLL - 1), prev = w == 0 ? W - 1 : w - 1;
  M1TimeMajor<BIG> mem;
  mem.A = A; mem.B = B; mem.iw = iw; mem.ih = ih; mem.dv = m1_div_make(iw);
  for (int G = w; G < groups; G += W) {
    const int y = G * 32 + lane;
    const bool rowint = y > 0 && y < ih - 1;
    const bool memup = rowint && (lane == 0 || y == 1);                           // the row above is not the lane above
    m1_row_setup(mem, rowint ? y : (G << 5) + (lane > 0 ? lane : 1));             // (rows that do nothing get harmless values)
    const int gbase = mem.gbase, R = mem.R, ubase = mem.ubase, uR = mem.uR, ushift = mem.ushift;
    M1Row r;
    r.gleft = 0; r.gleftF = 0u; r.croot[0] = r.croot[1] = -1;
    unsigned d0 = 0, d1 = 0, d2 = 0;                                               // (label | M1_FINAL)
    int q0 = 0, q1 = 0, q2 = 0, q3 = 0;
    auto wait_for = [&](int t0) {                                                  // the group above has finished what steps t0 .. t0 + M1_POLL - 1 read
      const int need = (G - 1) * (S + 1) + min(t0 + slack + 32 * M1_SKEW, S);
      unsigned spins = 0;
      while (prog[prev] < need) {
        __nanosleep(sleepns);                                                       // (a step takes some hundred ns, progress is published every M1_POLL steps)
        if (++spins > (1u << 18)) { *err = 1; __trap(); }                          // (never: a stuck wavefront ends the process loudly instead of hanging the device)
      }
      asm volatile("fence.acq_rel.cta;" ::: "memory");
    };
    if (G > 0) wait_for(0);
    // element (x', y - 1) while this lane is at step t: ubase + ((t + (x' - x) + ushift) mod iw) * uR
    auto up_ahead = [&](int tmv, int ahead) { return ubase + mem.wrap(mem.wrap(tmv + mem.wrap(ahead)) + ushift) * uR; };
    const int tm0 = mem.wrap(M1_SKEW * lane);                                      // (t mod iw) at the step of this lane's x = 0
    if (memup) {
      q3 = A[up_ahead(tm0, 0)];
      if (1 < iw) q2 = A[up_ahead(tm0, 1)];
      if (2 < iw) q1 = A[up_ahead(tm0, 2)];
      if (3 < iw) q0 = A[up_ahead(tm0, 3)];
    }
    // records two steps ahead: fn = of the pixel of this step, fn1 = of the next
    unsigned fn = 0, fn1 = 0;
    if (rowint && lane == 0) { fn = F[gbase]; if (1 < iw) fn1 = F[gbase + R]; }    // x = 0 of lane 0 is step 0
    int x = -M1_SKEW * lane, p = y * iw + x;
    int tm = 0;                                                                    // t mod iw, the same for every lane
    bool act = rowint && x == 0;
    unsigned aupn = 0;                                                             // (the shuffle of the next step's operand is issued a step early)
    for (int t = 0; t < S; t++) {
      if (G > 0 && t > 0 && (t & (M1_POLL - 1)) == 0) wait_for(t);
      const unsigned aup = memup ? (unsigned)q3 : aupn;                            // (from memory: not known to be final)
      q3 = q2; q2 = q1; q1 = q0;
      if (memup && (unsigned)(x + 4) < (unsigned)iw) q0 = A[up_ahead(tm, 4)];
      unsigned fin = (unsigned)r.gleft | r.gleftF;
      if (act) {
        const unsigned f = fn;
        mem.tm = tm;
        if (!(f & M1_INT)) { r.gleft = A[mem.self()]; r.gleftF = 0u; }
        else fin = m1_pixel(p, iw, f, aup, mem, r);
      }
      tm = tm + 1 == iw ? 0 : tm + 1;
      x++; p++;
      act = rowint && (unsigned)x < (unsigned)iw;
      fn = fn1;
      fn1 = (rowint && (unsigned)(x + 1) < (unsigned)iw) ? F[gbase + mem.wrap(tm + 1) * R] : 0u;   // the record of the step after the next
      d2 = d1; d1 = d0; d0 = fin;
      aupn = __shfl_up_sync(0xffffffffu, d2, 1);                                   // what the lane above produced three steps before the next one
      __syncwarp();
      if ((t & (M1_POLL - 1)) == M1_POLL - 1) {
        asm volatile("fence.acq_rel.cta;" ::: "memory");
        if (lane == 0) prog[w] = G * (S + 1) + t + 1;
      }
    }
  }
}
// label after the first pass = min(A, B), in the image layout
__global__ void k_m1_fold(int *out, const int *A, const int *B, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, A, B);
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  const int i = m1_index_xy(x, y, iw, ih);
  out[y * iw + x] = min(A[i], B[i]);
}
// the pointers of the first pass into the union-find of the pair labelling.  label[] holds tile roots that point at final roots
// (k_ccl_roots), so the component of a pixel is label[label[p]]; nearly every pointer stays inside its pixel's component, and along
// a row most of the rest repeat the union the lane to the left already asks for.
__global__ void k_merge_seed(int *label, const int *first, int n, size_t fs) {
  rd_batch_y(fs, label, first);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  int a = -1, b = -1;
  if (p < n) {
    const int t = first[p];
    if (t != p) { a = __ldcg(label + __ldcg(label + p)); b = __ldcg(label + __ldcg(label + t)); }
  }
  const int pa = __shfl_up_sync(0xffffffffu, a, 1), pb = __shfl_up_sync(0xffffffffu, b, 1);
  if (a != b && !((threadIdx.x & 31) > 0 && pa == a && pb == b)) rd_uf_unite(label, a, b);
}
// 0: the fixed point from the preprocess pointers (default); 1: the first pass replayed exactly, then the fixed point.  Process-wide;
// read when a labelling is enqueued (rd_rect.cu re-captures the CUDA graph of a page when the mode has changed).
static int g_merge_replay = -1;
extern "C" void rd_set_merge_replay(int on) { g_merge_replay = on != 0; }
extern "C" int rd_get_merge_replay(void) {
  if (g_merge_replay < 0) { const char *e = getenv("RD_MERGE_REPLAY"); g_merge_replay = e && atoi(e) != 0; }
  return g_merge_replay;
}
template <class MASK>
static void merge_core(int *out, int *work, LinkMerge<MASK> f, void *scratch, int *flags, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  // flags: RD_MERGE_ROUNDS round flags + the guard word of the wavefront
  RD_CUDA(cudaMemset2DAsync(flags, fs ? fs : (RD_MERGE_ROUNDS + 1) * sizeof(int), 0, (RD_MERGE_ROUNDS + 1) * sizeof(int), nb, s));
  const int n = iw * ih, g4 = rd_cdiv(rd_cdiv(n, 4), 256);
  const dim3 b(32, 8);
  const bool replay = rd_get_merge_replay() != 0;
  f.pre = replay ? 0 : 1;
  if (replay) {
    // the first pass: A = work, B = scratch (as a plane of ints), records in `out` - whose image-layout content the fold then writes
    RD_LAUNCH(k_m1_pre<MASK>, rd_gz(rd_grid2d(iw, ih, b), nb), b, 0, s, work, (int *)scratch, (uint8_t *)out, f, iw, ih, fs);
    // warps per CTA: a group of 32 rows starts 32 * M1_SKEW + 2 * M1_POLL steps after the one above and takes iw + 31 * M1_SKEW steps,
    // so no more than that many groups (+ 1) are ever under way at once; further warps would only hold registers
    // cushion between two groups beyond the 32 * M1_SKEW steps the data flow needs: the smallest the polling allows.  (Measured: the
    // kernel time is the length of the chain, (groups - 1) * (128 + cushion) + iw + 124 steps, times 0.61 - 0.65 us whatever the
    // cushion - 2.96 / 3.35 / 3.95 / 4.72 ms at 720p for 16 / 48 / 96 / 160: the warps do not hold each other up, a step is simply slow.)
    const int slack = 2 * M1_POLL;
    const int sleepns = 1500;                     // (measured: 100 ... 5000 ns make no difference, the waiting warps are not on the critical path)
    const int groups = rd_cdiv(ih, 32), live = (iw + M1_SKEW * 31) / (32 * M1_SKEW + slack) + 2;
    const int wv = groups < live ? groups : (live < 32 ? live : 32);
    if (iw >= M1_BIG && n < (1 << 24))
      RD_LAUNCH(k_m1_wave<true>, nb, 32 * wv, 0, s, work, (int *)scratch, (const uint8_t *)out, flags + RD_MERGE_ROUNDS, iw, ih, slack, sleepns, fs);
    else
      RD_LAUNCH(k_m1_wave<false>, nb, 32 * wv, 0, s, work, (int *)scratch, (const uint8_t *)out, flags + RD_MERGE_ROUNDS, iw, ih, slack, sleepns, fs);
    RD_LAUNCH(k_m1_fold, rd_gz(rd_grid2d(iw, ih, b), nb), b, 0, s, out, (const int *)work, (const int *)scratch, iw, ih, fs);
  }
  // pairs that may adopt in both directions (+ the preprocess pointers, or afterwards the pointers of the first pass)
  ccl_core(work, (uint8_t *)scratch, f, iw, ih, nb, fs, s);
  if (replay) {
    RD_LAUNCH(k_merge_seed, rd_gy(rd_cdiv(n, 256), nb), 256, 0, s, work, (const int *)out, n, fs);
    RD_LAUNCH(k_ccl_roots, rd_gy(g4, nb), 256, 0, s, work, (const uint8_t *)scratch, n, fs);
  }
  // round 0 has work on a third of the frames, the kernels behind the gate of a later round practically never: small grids there
  // (grid-stride loops)
  const int gs = g4 < 48 ? g4 : 48;
  for (int r = 0; r < RD_MERGE_ROUNDS; r++) {
    RD_LAUNCH(k_merge_gate<MASK>, rd_gy(g4, nb), 256, 0, s, (const int *)work, (uint8_t *)scratch, f, flags, r, iw, ih, fs);
    RD_LAUNCH(k_merge_apply, rd_gy(r == 0 ? g4 : gs, nb), 256, 0, s, work, (uint8_t *)scratch, (const int *)flags, r, iw, n, fs);
    RD_LAUNCH(k_merge_roots, rd_gy(r == 0 ? g4 : gs, nb), 256, 0, s, work, (const uint8_t *)scratch, (const int *)flags, r, n, fs);
  }
  if (replay) RD_LAUNCH(k_ccl_flatten_merge1, rd_gz(rd_grid2d(iw, ih, b), nb), b, 0, s, out, (const int *)work, iw, ih, fs);
  else {
    RD_LAUNCH(k_ccl_flatten_merge, rd_gz(rd_grid2d(iw, ih, b), nb), b, 0, s, out, work, f.pix, iw, ih, fs);
    RD_LAUNCH(k_merge_toprow, nb, 32, 0, s, out, f.pix, f.edge, iw, ih, fs);
  }
}
// labelxPreprocess + 8 x labelMergeMain (oclrect.c:325-331).  work: iw*ih ints, scratch: iw*ih bytes (iw*ih ints with the first-pass replay), flags: RD_MERGE_ROUNDS + 1 ints;
// out may not alias work
void rd_labelMerge(int *out, int *work, const uint32_t *pix, const int *mask, const int *edge, void *scratch, int *flags, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  LinkMerge<int> f = {pix, mask, edge, iw, ih, 1};
  merge_core(out, work, f, scratch, flags, iw, ih, nb, fs, s);
}
// same with the merge mask as a byte plane
void rd_labelMerge_u8(int *out, int *work, const uint32_t *pix, const uint8_t *mask, const int *edge, void *scratch, int *flags, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  LinkMerge<uint8_t> f = {pix, mask, edge, iw, ih, 1};
  merge_core(out, work, f, scratch, flags, iw, ih, nb, fs, s);
}
