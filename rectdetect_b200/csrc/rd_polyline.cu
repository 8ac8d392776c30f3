// rd_polyline.cu - Stage C: edge-string tracing and polyline simplification (oclpolyline.h:85-88) on sm_100a.
//
// oclpolyline_execute keeps the reference's contract (oclpolyline.c:218-309): `in` is a 0/1 edge mask, the results
// are the per-pixel segment-id map `lsIdOut` and the linesegment_t list `lsList` (element 0 = header).  The kernels
// follow oclpolyline.cl one for one, except where the reference is schedule dependent; there the canonical outcome
// of DESIGN.md is computed deterministically:
//   - the three bounded label-propagation loops (label8x, labelpl) are exact connected components (rd_ccl.cu);
//   - relabel_pass0 numbers strings in raster order of their root pixel by a block-count / scan / rank sequence
//     instead of atomic_inc arrival order (Q4);
//   - mkpl_pass2 picks the arg-max pixel with the smallest index and hands out new ids by a prefix sum over the
//     parent id (Q4, Q8); because every split only rewrites fields of its own, its new and its right neighbour's
//     entry, the 15 full-list copies of the reference (oclpolyline.c:207, 32 B/px each) are not needed at all;
//   - refine_pass3 computes all shared vertices from the unmodified list before writing any (Q5).
#include "rd_common.cuh"
#include "rd_stageA.cuh"
#include "rd_bits.cuh"
#include <cooperative_groups.h>

struct oclpolyline_t { uint32_t magic; int ordinal; };
#define POLY_MAGIC 0x808eae03u

typedef linesegment_t LS_t;
struct LSX_t {                         // oclpolyline.cl:41-45
  long long mx00, mx01, mx11, my0, my1;
  short dirSEx, dirSEy, vDirSEx, vDirSEy;
  int distSquSE, padding;
};
static_assert(sizeof(LS_t) == 56 && sizeof(LSX_t) == 56, "list entries are 56 bytes");

#define MINEDGELEN 1
#define MINNINDEX 4
#define XY2D const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y; if (x >= iw || y >= ih) return; const int p0 = y * iw + x
#define IS_BORDER1 (x <= 0 || y <= 0 || x >= iw - 1 || y >= ih - 1)
// kernels over the compact list of labelled pixels: grid-stride loop, `continue` instead of `return`
// The list kernel (kp_polyline_list) runs one thread-block CLUSTER per frame: its passes are grid-stride loops over the
// cluster's threads, separated by cluster barriers (hardware barrier.cluster with release / acquire ordering of the global
// lists they exchange).
#define PL_CLUSTER 8
#define PL_THREADS 512
__device__ __forceinline__ int pl_tid() { return (int)(cooperative_groups::this_cluster().block_rank() * blockDim.x + threadIdx.x); }
__device__ __forceinline__ int pl_nt() { return (int)(cooperative_groups::this_cluster().num_blocks() * blockDim.x); }
#define PLIST_LOOP const int pcount_ = plist[0]; for (int k_ = pl_tid(), nt_ = pl_nt(); k_ < pcount_; k_ += nt_)
#define PL_SEG_LOOP(count) for (int g = pl_tid() + 1, nt_ = pl_nt(); g <= (count); g += nt_)
#define PLIST_XY const int p0 = plist[k_ + 1]; const int x = p0 % iw, y = p0 / iw; (void)x; (void)y
#define LIST_BLOCKS 32

void rd_label8x(int *label, const int *pix, void *scratch, int bgc, int iw, int ih, int nb, size_t fs, cudaStream_t s);
void rd_labelpl(int *label, const int *num, void *scratch, int iw, int ih, int nb, size_t fs, cudaStream_t s);
void rd_k_clear(int *out, int nints, int nb, size_t fs, cudaStream_t s);
void rd_k_copy(int *out, const int *in, int nints, int nb, size_t fs, cudaStream_t s);
void rd_k_rand(int *out, uint64_t seed, int n, int nb, size_t fs, cudaStream_t s);

// ---------------------------------------------------------------------------- string clean-up (oclpolyline.cl:66-147)
__global__ void kp_simpleJunction(int *out, const int *in, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, in);
  XY2D;
  int r = 0;
  if (!IS_BORDER1 && in[p0] != 0) {
    int count = 1;
#pragma unroll
    for (int i = 0; i < 8; i++) if (in[p0 + RD_RX[i] + RD_RY[i] * iw] != 0) count++;
    r = count == 1 ? 0 : count;
  }
  out[p0] = r;
}
// the 2-pixel frame of `out` is left untouched, as in the reference (oclpolyline.cl:91)
__global__ void kp_simpleConnect(int *out, const int *in, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, in);
  XY2D;
  if (x <= 1 || y <= 1 || x >= iw - 2 || y >= ih - 2) return;
  int r = in[p0] != 0 ? 1 : 0;
  if (!r) {
    if (in[p0 - 2] != 0 && in[p0 - 1] == 2 && in[p0 + 1] == 2 && in[p0 + 2] != 0) r = 1;
    if (in[p0 - iw * 2] != 0 && in[p0 - iw] == 2 && in[p0 + iw] == 2 && in[p0 + iw * 2] != 0) r = 1;
    if (in[p0 - iw * 2 - 2] != 0 && in[p0 - iw - 1] == 2 && in[p0 + iw + 1] == 2 && in[p0 + iw * 2 + 2] != 0) r = 1;
    if (in[p0 - iw * 2 + 2] != 0 && in[p0 - iw + 1] == 2 && in[p0 + iw - 1] == 2 && in[p0 + iw * 2 - 2] != 0) r = 1;
    if (in[p0 + 2] != 0 && in[p0 + 1] == 2 && in[p0 + iw - 1] == 2 && in[p0 + iw - 2] != 0) r = 1;
    if (in[p0 - 2] != 0 && in[p0 - 1] == 2 && in[p0 + iw + 1] == 2 && in[p0 + iw + 2] != 0) r = 1;
    if (in[p0 - iw * 2 + 1] != 0 && in[p0 - iw + 1] == 2 && in[p0 + iw] == 2 && in[p0 + iw * 2] != 0) r = 1;
    if (in[p0 - iw * 2 - 1] != 0 && in[p0 - iw - 1] == 2 && in[p0 + iw] == 2 && in[p0 + iw * 2] != 0) r = 1;
  }
  out[p0] = r;
}
__global__ void kp_stringify(int *out, const int *in, int mod2, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, in);
  XY2D;
  int r = in[p0];
  if (!IS_BORDER1 && ((x + y) & 1) == mod2) {
    const bool n = in[p0 - iw] != 0, s = in[p0 + iw] != 0, w = in[p0 - 1] != 0, e = in[p0 + 1] != 0;
    if ((n || s) && (w || e)) r = 0;
  }
  out[p0] = r;
}
__global__ void kp_removeBranch(int *out, const int *in, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, in);
  XY2D;
  int r = 0;
  if (!IS_BORDER1 && in[p0] != 0) {
    int count = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) if (in[p0 + RD_RX[i] + RD_RY[i] * iw] != 0) count++;
    r = count <= 2 ? 1 : 0;
  }
  out[p0] = r;
}
// oclpolyline.cl:149-167 ; the reference's non-atomic ++ is only ever compared with 0 (Q7)
__global__ void kp_countEnds(int *out, const int *junction, const int *label, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, junction, label);
  XY2D;
  if (IS_BORDER1) return;
  if (junction[p0] == 2) out[label[p0]] = 1;
}
__global__ void kp_breakLoops(int *edgeinout, int *labelinout, const int *nEnds, int iw, int ih, size_t fs) {
  rd_batch_z(fs, edgeinout, labelinout, nEnds);
  XY2D;
  if (IS_BORDER1) return;
  if (labelinout[p0] != p0) return;
  if (nEnds[p0] == 0) { edgeinout[p0] = 0; labelinout[p0] = -1; }
}

// ---------------------------------------------------------------------------- end finding / numbering (oclpolyline.cl:169-310)
__device__ __forceinline__ void getnp(const int *labelin, int p0, int iw, int &nx, int &ny) {
  const int l = labelin[p0];
  int i;
  for (i = 0; i < 8; i++) if (labelin[p0 + RD_RX[i] + RD_RY[i] * iw] == l) break;
  nx = i < 8 ? (p0 + RD_RX[i] + RD_RY[i] * iw) : p0;
  for (i++; i < 8; i++) if (labelin[p0 + RD_RX[i] + RD_RY[i] * iw] == l) break;
  ny = i < 8 ? (p0 + RD_RX[i] + RD_RY[i] * iw) : p0;
}
__global__ void kp_findEnds0(int *nextout, int *prevout, int *flagout, const int *labelin, int iw, int ih, size_t fs) {
  rd_batch_z(fs, nextout, prevout, flagout, labelin);
  XY2D;
  int nx = -1, pv = -1, flag = -1;
  if (!IS_BORDER1 && labelin[p0] != -1) {
    int npx, npy, a, b;
    getnp(labelin, p0, iw, npx, npy);
    nx = npx; pv = npy; flag = 0;
    if (npx != p0) { getnp(labelin, npx, iw, a, b); if (a == p0) flag |= 1; }
    if (npy != p0) { getnp(labelin, npy, iw, a, b); if (b == p0) flag |= 2; }
  }
  nextout[p0] = nx; prevout[p0] = pv; flagout[p0] = flag;
}
// A launch reads one pair of flag bits of other pixels and rewrites only the other pair of its own pixel, so the
// in-place update of flaginout is race free (the reads are volatile to keep them from being cached in registers).
__global__ void kp_findEnds1(int *nextout, int *prevout, int *flaginout, const int *nextin, const int *previn, const int *labelin, int page, int iw, int ih, size_t fs) {
  rd_batch_z(fs, nextout, prevout, flaginout, nextin, previn, labelin);
  XY2D;
  int nn = -1, pp = -1;
  if (!IS_BORDER1 && labelin[p0] != -1) {
    const volatile int *fl = flaginout;
    const int f0 = fl[p0];
    bool revn = page == 0 ? ((f0 & 1) != 0) : ((f0 & 4) != 0);
    bool revp = page == 0 ? ((f0 & 2) != 0) : ((f0 & 8) != 0);
    nn = nextin[p0]; pp = previn[p0];
    for (int i = 0; i < 8; i++) {
      const int nn2 = revn ? previn[nn] : nextin[nn];
      const int pp2 = revp ? nextin[pp] : previn[pp];
      int nflag = fl[nn], pflag = fl[pp];
      if (page != 0) { nflag >>= 2; pflag >>= 2; }
      revn = revn ? ((nflag & 2) == 0) : ((nflag & 1) != 0);
      revp = revp ? ((pflag & 1) == 0) : ((pflag & 2) != 0);
      nn = nn2; pp = pp2;
    }
    int f = f0;
    if (page == 0) { f &= 3; f |= revn ? 4 : 0; f |= revp ? 8 : 0; }
    else { f &= (3 << 2); f |= revn ? 1 : 0; f |= revp ? 2 : 0; }
    flaginout[p0] = f;
  }
  nextout[p0] = nn; prevout[p0] = pp;
}
__global__ void kp_findEnds2(int *numout, int *linkout, const int *nextin, const int *previn, const int *labelin, int iw, int ih, size_t fs) {
  rd_batch_z(fs, numout, linkout, nextin, previn, labelin);
  XY2D;
  int num = 0, link = -1;
  if (!IS_BORDER1 && labelin[p0] != -1) {
    int npx, npy;
    getnp(labelin, p0, iw, npx, npy);
    link = nextin[p0] < previn[p0] ? npx : npy;
    num = link == p0 ? 0 : 1;
  }
  numout[p0] = num; linkout[p0] = link;
}
__global__ void kp_number(int *numout, int *linkout, const int *numin, const int *linkin, int iw, int ih, size_t fs) {
  rd_batch_z(fs, numout, linkout, numin, linkin);
  XY2D;
  int no = 0, lo = -1;
  if (!IS_BORDER1) {
    const int l0 = linkin[p0];
    if (l0 == -1) { no = numin[p0]; lo = -1; }
    else {
      int n = numin[p0], l = l0;
      bool bail = false;
      for (int i = 0; i < 32; i++) {
        if (!(0 < l && l < iw * ih)) { bail = true; break; }
        n += numin[l];
        l = linkin[l];
      }
      if (!bail) { no = n; lo = l; }
    }
  }
  numout[p0] = no; linkout[p0] = lo;
}

// ---------------------------------------------------------------------------- labelpl pre-step, sizes (oclpolyline.cl:312-378)
__global__ void kp_plus1(int *out, const int *in, int n, size_t fs) {
  rd_batch_y(fs, out, in);          // labelpl_preprocess: pix = number + 1 where number != 0
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const int v = in[i]; out[i] = v == 0 ? 0 : v + 1; }
}
__global__ void kp_calcSize(int *out, const int *label, int n, size_t fs) {
  rd_batch_y(fs, out, label);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = label[i];
  if (b != 0) atomicAdd(out + b, 1);
}
__global__ void kp_filterSize(int *out, const int *labelin, const int *sizein, int sizethre, int n, size_t fs) {
  rd_batch_y(fs, out, labelin, sizein);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = labelin[i];
  out[i] = sizein[b] > sizethre ? b : 0;
}

// ---------------------------------------------------------------------------- relabel (oclpolyline.cl:380-420), raster-order ids
// A root is an interior pixel whose label is its own index.  ids = 1 + number of roots before it in raster order.
#define RL_BLOCK 1024
__device__ __forceinline__ bool is_root(const int *label, int p, int iw, int ih) {
  const int x = p % iw, y = p / iw;
  if (x <= 0 || y <= 0 || x >= iw - 1 || y >= ih - 1) return false;
  const int g = label[p];
  return g != 0 && g == p;
}
__global__ void kp_relabel_count(int *blockCount, const int *label, int iw, int ih, size_t fs) {
  rd_batch_y(fs, blockCount, label);
  __shared__ int wsum[RL_BLOCK / 32];
  const int p = blockIdx.x * RL_BLOCK + threadIdx.x;
  const bool r = p < iw * ih && is_root(label, p, iw, ih);
  const unsigned b = __ballot_sync(0xffffffffu, r);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = __popc(b);
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = wsum[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) blockCount[blockIdx.x] = v;
  }
}
// exclusive scan of blockCount in place (single CTA, chunked with a running carry); total -> *total
__global__ void kp_scan_blocks(int *blockCount, int nblocks, int *total, int *plist, size_t fs) {
  rd_batch_x(fs, blockCount, total, plist);
  if (threadIdx.x == 0) plist[0] = 0;
  __shared__ int wsum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nblocks ? blockCount[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = wsum[threadIdx.x], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, wi, o); if (threadIdx.x >= o) wi += t; }
      wsum[threadIdx.x] = wi - w;
    }
    __syncthreads();
    const int excl = carry + wsum[threadIdx.x >> 5] + incl - v;
    if (i < nblocks) blockCount[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void kp_relabel_rank(int *table, const int *blockOffset, const int *label, int iw, int ih, size_t fs) {
  rd_batch_y(fs, table, blockOffset, label);
  __shared__ int wsum[RL_BLOCK / 32];
  const int p = blockIdx.x * RL_BLOCK + threadIdx.x;
  const bool r = p < iw * ih && is_root(label, p, iw, ih);
  const unsigned b = __ballot_sync(0xffffffffu, r);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) wsum[w] = __popc(b);
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = wsum[threadIdx.x], vi = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, vi, o); if (threadIdx.x >= o) vi += t; }
    wsum[threadIdx.x] = vi - v;
  }
  __syncthreads();
  if (r) table[p + 1] = blockOffset[blockIdx.x] + wsum[w] + __popc(b & ((1u << lane) - 1)) + 1;
}
// relabel_pass1 (oclpolyline.cl:400) + compaction: every pixel that ends up with a segment id is appended to plist
// (plist[0] = count, entries from 1; order is irrelevant to the consumers, which only use commutative atomics).
__global__ void kp_relabel_pass1(int *labelinout, const int *tablein, int *plist, int iw, int ih, size_t fs) {
  rd_batch_z(fs, labelinout, tablein, plist);
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  const bool in = x < iw && y < ih;
  const int p0 = y * iw + x;
  int g = 0;
  if (in) {
    if (x == 0 || y == 0 || x >= iw - 1 || y >= ih - 1) labelinout[p0] = 0;
    else {
      g = labelinout[p0];
      if (g != 0) { g = tablein[g + 1]; labelinout[p0] = g; }
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, g != 0);
  if (m == 0) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(plist, __popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (g != 0) plist[1 + base + __popc(m & ((1u << lane) - 1))] = p0;
}

// ---------------------------------------------------------------------------- mkpl (oclpolyline.cl:439-646)
__device__ __forceinline__ bool ls_overflow(int g, int lsListSize) { return g < 0 || (size_t)lsListSize <= ((size_t)(g + 1)) * sizeof(LS_t); }

__device__ __forceinline__ float distanceSqu(float vx, float vy, float wx, float wy) {
  const float dx = __fsub_rn(vx, wx), dy = __fsub_rn(vy, wy);
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}
__device__ __forceinline__ void closestPoint(float vx, float vy, float wx, float wy, float px, float py, float &ox, float &oy) {
  const float l2 = distanceSqu(vx, vy, wx, wy);
  if (l2 <= 1e-4f) { ox = vx; oy = vy; return; }
  const float t = __fdiv_rn(__fadd_rn(__fmul_rn(__fsub_rn(px, vx), __fsub_rn(wx, vx)), __fmul_rn(__fsub_rn(py, vy), __fsub_rn(wy, vy))), l2);
  if (t < 0.0f) { ox = vx; oy = vy; return; }
  if (t > 1.0f) { ox = wx; oy = wy; return; }
  ox = __fadd_rn(vx, __fmul_rn(t, __fsub_rn(wx, vx)));
  oy = __fadd_rn(vy, __fmul_rn(t, __fsub_rn(wy, vy)));
}

// init: zero the list entries of the K initial strings (K = table[0], oclpolyline.c:192 clears the whole list), reset
// the per-string start / end pixel slots and the arg-max slots, and arm the iteration flags (oclpolyline.cl:442-444)
__device__ __forceinline__ void d_mkpl_init(LS_t *gp, int lsListSize, int *aux, int *winner, int cap, const int *table, int *flags, int maxIter) {
  const int cap2 = min(cap, lsListSize / (int)sizeof(LS_t));
  const int K = min(table[0], cap2 - 1);
  const int t = pl_tid(), nt = pl_nt();
  if (t < maxIter + 1) flags[t] = t == 0 ? 1 : 0;
  int *raw = (int *)gp;
  for (int i = t; i < (K + 1) * 14; i += nt) raw[i] = 0;
  for (int g = t; g <= K; g += nt) { aux[g] = 0; aux[cap + g] = 0x7fffffff; winner[g] = 0x7fffffff; }
}
// pass0a: per-string start pixel (the LAST pixel in raster order whose number is 1, as a sequential sweep leaves it),
// pixel count, largest number, and the list header (largest id).  aux[g] / aux[cap+g] : start / end pixel index.
__device__ __forceinline__ void d_mkpl_pass0a(LS_t *gp, int lsListSize, int *aux, int cap, const int *numberin, const int *labelin, const int *plist, int iw) {
  PLIST_LOOP {
    PLIST_XY;
    const int g = labelin[p0], n = numberin[p0];
    if (g == 0 || ls_overflow(g, lsListSize)) continue;
    if (n == 1) { atomicMax(aux + g, p0 + 1); atomicAdd(&gp[g].startCount, 1); }
    atomicAdd(&gp[g].npix, 1);
    atomicMax(&gp[g].endIndex, n);
    atomicMax((int *)gp, g);
  }
}
__device__ __forceinline__ void d_mkpl_pass0b(LS_t *gp, int lsListSize, int *aux, int cap, const int *numberin, const int *labelin, const int *plist, int iw) {
  PLIST_LOOP {
    PLIST_XY;
    const int g = labelin[p0], n = numberin[p0];
    if (g == 0 || ls_overflow(g, lsListSize)) continue;
    if (n == gp[g].endIndex) {
      if (gp[g].startCount == 1 && gp[g].npix >= 2) { atomicAdd(&gp[g].endCount, 1); atomicMin(aux + cap + g, p0); }
      else aux[cap + g] = -1;                                      // polyid = 0
    }
  }
}
// one thread per list entry: turn the start / end pixel indices into coordinates
__device__ __forceinline__ void d_mkpl_pass0c(LS_t *gp, const int *aux, int cap, int iw) {
  const int count = min(*(const int *)gp, cap - 1);
  PL_SEG_LOOP(count) {
    const int sp = aux[g] - 1, ep = aux[cap + g];
    if (sp >= 0) { gp[g].x0 = (float)(sp % iw); gp[g].y0 = (float)(sp / iw); gp[g].level = 0; }
    if (ep >= 0 && ep != 0x7fffffff) { gp[g].x1 = (float)(ep % iw); gp[g].y1 = (float)(ep / iw); gp[g].polyid = g; }
    else gp[g].polyid = 0;
  }
}
__global__ void kp_fill(int *out, int v, int n, size_t fs) {
  rd_batch_y(fs, out);
  const int i = (int)threadIdx.x;
  if (i < n) out[i] = v;
}

// pass1 (oclpolyline.cl:509): distance of every labelled pixel to the chord of its segment, xor-ed with 13 bits of the
// per-pixel hash (oclpolyline.cl:883, seed 0, evaluated in place instead of being read from a plane), kept per list
// slot, and the per-segment maximum.
__device__ __forceinline__ void d_mkpl_pass1(LS_t *gp, int lsListSize, int *dist, const int *labelin, const int *plist, const int *flags, int nIter, int iw) {
  PLIST_LOOP {
    PLIST_XY;
    const int g = labelin[p0];
    if (g == 0 || ls_overflow(g, lsListSize)) continue;
    if (gp[g].polyid == 0) continue;
    const int x0 = (int)gp[g].x0, y0 = (int)gp[g].y0, x1 = (int)gp[g].x1, y1 = (int)gp[g].y1;
    float cx, cy;
    closestPoint((float)x0, (float)y0, (float)x1, (float)y1, (float)x, (float)y, cx, cy);
    int d = (int)__fmul_rn(rd_hypot(__fsub_rn(cx, (float)x), __fsub_rn(cy, (float)y)), 65536.0f);
    d ^= (rd_rand_at(p0, 0) & 0x1fff);
    dist[k_] = d;
    atomicMax(&gp[g].maxDist, d);
  }
}
// pass2a: winner[g] = smallest pixel index attaining maxDist
__device__ __forceinline__ void d_mkpl_pass2a(const LS_t *gp, int lsListSize, int *winner, const int *dist, const int *labelin, const int *plist, const int *flags, int nIter, int iw) {
  PLIST_LOOP {
    PLIST_XY;
    const int g = labelin[p0];
    if (g == 0 || ls_overflow(g, lsListSize)) continue;
    if (g > *(const int *)gp) continue;
    if (gp[g].polyid == 0) continue;
    if (dist[k_] != gp[g].maxDist) continue;
    atomicMin(winner + g, p0);
  }
}
// pass2b: ONE CTA (rank 0 of the cluster); decides the splits of this iteration, numbers the new entries in raster order of
// the splitting pixels (the reference's atomic_inc arrival order when its work-items run in raster order, oclpolyline.cl:585)
// and rewrites the list.  Also resets winner[] for the next iteration.
#ifndef PL_SPLITCAP
#define PL_SPLITCAP 2048   // splitting pixels of one round ranked out of shared memory; more than that: ranked against winner[] in global memory
#endif
__device__ __forceinline__ bool d_mkpl_splits(const LS_t &o, int px, int py, float minerror) {
  if (o.endIndex - o.startIndex < MINNINDEX - 1) return false;
  if (o.startCount > 1 || o.endCount > 1) return false;
  const int maxDist = o.maxDist;
  if (maxDist < ((int)__fmul_rn(minerror, 65536.0f))) return false;
  if ((float)maxDist < __fmul_rn(__fmul_rn(minerror, 3.0f), 65536.0f) &&
      __fdiv_rn(__fmul_rn((float)maxDist, (float)maxDist), distanceSqu(o.x0, o.y0, o.x1, o.y1)) < 100000.0f) return false;
  if (distanceSqu((float)px, (float)py, o.x0, o.y0) < (float)(MINEDGELEN * MINEDGELEN)) return false;
  if (distanceSqu((float)px, (float)py, o.x1, o.y1) < (float)(MINEDGELEN * MINEDGELEN)) return false;
  return true;
}
__device__ __forceinline__ void d_mkpl_pass2b(LS_t *gp, int lsListSize, int *winner, const int *numberin, const int *flags, int nIter, float minerror, int iw) {
  __shared__ int spix[PL_SPLITCAP];
  __shared__ int nsplit;
  const int count = *(const int *)gp;
  if (threadIdx.x == 0) nsplit = 0;
  __syncthreads();
  // phase A: winner[g] keeps the splitting pixel of the segments that split, "none" otherwise
  for (int base = 1; base <= count; base += PL_THREADS) {
    const int g = base + threadIdx.x;
    if (g > count || ls_overflow(g, lsListSize)) continue;
    const int p0 = winner[g];
    bool split = false;
    if (p0 != 0x7fffffff) {
      const LS_t o = gp[g];
      split = o.polyid != 0 && d_mkpl_splits(o, p0 % iw, p0 / iw, minerror);
    }
    if (split) {
      const int k = atomicAdd(&nsplit, 1);
      if (k < PL_SPLITCAP) spix[k] = p0;
    } else winner[g] = 0x7fffffff;
  }
  __syncthreads();
  const int S = nsplit;
  // phase B: id of a new entry = count + 1 + number of splitting pixels in front of its own in raster order
  for (int base = 1; base <= count; base += PL_THREADS) {
    const int g = base + threadIdx.x;
    if (g > count || ls_overflow(g, lsListSize)) continue;
    const int p0 = winner[g];
    if (p0 == 0x7fffffff) continue;
    int rank = 0;
    if (S <= PL_SPLITCAP) { for (int k = 0; k < S; k++) rank += spix[k] < p0; }
    else { for (int h = 1; h <= count; h++) rank += winner[h] < p0; }
    const int gn = count + rank + 1;
    if (ls_overflow(gn, lsListSize)) continue;
    const LS_t o = gp[g];
    const int px = p0 % iw, py = p0 / iw, n = numberin[p0], gr = o.rightPtr;
    LS_t nw;
    nw.x0 = (float)px; nw.y0 = (float)py; nw.x1 = o.x1; nw.y1 = o.y1;
    nw.startIndex = n; nw.endIndex = o.endIndex; nw.leftPtr = g; nw.rightPtr = gr;
    nw.startCount = 0; nw.endCount = 0; nw.maxDist = 0; nw.polyid = o.polyid; nw.npix = 0; nw.level = o.maxDist;
    gp[gn] = nw;
    winner[gn] = 0x7fffffff;
    gp[g].endIndex = n; gp[g].x1 = (float)px; gp[g].y1 = (float)py; gp[g].rightPtr = gn; gp[g].maxDist = 0;
    if (gr != 0) gp[gr].leftPtr = gn;
  }
  __syncthreads();
  for (int g = 1 + threadIdx.x; g <= count; g += PL_THREADS) if (!ls_overflow(g, lsListSize)) winner[g] = 0x7fffffff;
  if (threadIdx.x == 0) *(int *)gp = count + S;
}
__device__ __forceinline__ void d_mkpl_pass3(const LS_t *gp, int lsListSize, const int *numberin, int *labelinout, const int *plist, int *flags, int nIter, int iw) {
  PLIST_LOOP {
    PLIST_XY;
    const int g = labelinout[p0];
    if (g == 0 || ls_overflow(g, lsListSize)) continue;
    if (gp[g].polyid == 0) continue;
    if (gp[g].endIndex < numberin[p0]) { labelinout[p0] = gp[g].rightPtr; flags[nIter] = 1; }
  }
}

// ---------------------------------------------------------------------------- refine (oclpolyline.cl:680-809)
__device__ __forceinline__ void d_refine_pass0(LSX_t *lsx, const LS_t *ls) {
  const int count_ = *(const int *)ls;
  PL_SEG_LOOP(count_) {
  if (ls[g].polyid == 0) continue;
  LSX_t v;
  v.dirSEx = (short)(int)__fsub_rn(ls[g].x1, ls[g].x0);          // convert_short2: truncation (Q18)
  v.dirSEy = (short)(int)__fsub_rn(ls[g].y1, ls[g].y0);
  v.vDirSEx = (short)(-v.dirSEy);
  v.vDirSEy = v.dirSEx;
  v.mx00 = v.mx01 = v.mx11 = v.my0 = v.my1 = 0;
  v.distSquSE = v.dirSEx * v.dirSEx + v.dirSEy * v.dirSEy;
  v.padding = 0;
  lsx[g] = v;
  }
}
__device__ __forceinline__ void d_refine_pass1(LSX_t *lsx, const LS_t *ls, const int *lsIdIn, const int *plist, int iw) {
  PLIST_LOOP {
    PLIST_XY;
    const int g = lsIdIn[p0];
    if (g == 0) continue;
    if (g < 0 || *(const int *)ls < g) continue;
    const int vx = x - __float2int_rn(ls[g].x0), vy = y - __float2int_rn(ls[g].y0);   // convert_int2_rte
    const int ay = vx * (int)lsx[g].vDirSEx + vy * (int)lsx[g].vDirSEy;
    const int ax0 = vx * (int)lsx[g].dirSEx + vy * (int)lsx[g].dirSEy;
    const int ax1 = lsx[g].distSquSE;
    typedef unsigned long long ull;
    atomicAdd((ull *)&lsx[g].mx00, (ull)__float2ll_rn(__fmul_rn((float)ax0, (float)ax0)));   // convert_long_rte
    atomicAdd((ull *)&lsx[g].mx01, (ull)__float2ll_rn(__fmul_rn((float)ax0, (float)ax1)));
    atomicAdd((ull *)&lsx[g].mx11, (ull)__float2ll_rn(__fmul_rn((float)ax1, (float)ax1)));
    atomicAdd((ull *)&lsx[g].my0, (ull)__float2ll_rn(__fmul_rn((float)ax0, (float)ay)));
    atomicAdd((ull *)&lsx[g].my1, (ull)__float2ll_rn(__fmul_rn((float)ax1, (float)ay)));
  }
}
__device__ __forceinline__ void d_refine_pass2(const LSX_t *lsx, LS_t *ls) {
  const int count_ = *(const int *)ls;
  PL_SEG_LOOP(count_) {
  if (ls[g].polyid == 0) continue;
  const float mx00 = (float)lsx[g].mx00, mx01 = (float)lsx[g].mx01, mx11 = (float)lsx[g].mx11, my0 = (float)lsx[g].my0, my1 = (float)lsx[g].my1;
  float rdet = __fsub_rn(__fmul_rn(mx00, mx11), __fmul_rn(mx01, mx01));
  if (rdet == 0) continue;
  rdet = (float)__ddiv_rn(1.0, (double)rdet);                   // `1.0 / rdet` divides in double (Q15)
  const float as0 = __fmul_rn(__fsub_rn(__fmul_rn(mx11, my0), __fmul_rn(mx01, my1)), rdet);
  const float as1 = __fmul_rn(__fsub_rn(__fmul_rn(mx00, my1), __fmul_rn(mx01, my0)), rdet);
  const float vx = (float)lsx[g].vDirSEx, vy = (float)lsx[g].vDirSEy, as01 = __fadd_rn(as0, as1);
  ls[g].x0 = __fadd_rn(ls[g].x0, __fmul_rn(vx, as1));
  ls[g].y0 = __fadd_rn(ls[g].y0, __fmul_rn(vy, as1));
  ls[g].x1 = __fadd_rn(ls[g].x1, __fmul_rn(vx, as01));
  ls[g].y1 = __fadd_rn(ls[g].y1, __fmul_rn(vy, as01));
  }
}
// refine_pass3 (oclpolyline.cl:772-809) rewrites the vertex a segment shares with its right neighbour IN PLACE: what a
// work-item reads depends on which of its neighbours ran before it.  Canonical (Q5) = the work-items in id order: g sees its
// own start as moved by its left neighbour l iff l < g, and the end of its right neighbour h as moved iff h < g (and h has a
// right neighbour itself).  Neighbouring segments are therefore always ordered by a dependency, dependencies point to smaller
// ids, and a chain of them is a run of increasing ids along a polyline, i.e. at most one per split round: ONE CTA evaluates
// the list level by level, in place (`state`: 0 = waiting, 1 = done).
__device__ __forceinline__ void d_refine_pass3(LS_t *ls, int *state) {
  __shared__ int pending;
  const int count = *(const int *)ls;
  for (int g = 1 + threadIdx.x; g <= count; g += blockDim.x) state[g] = (ls[g].polyid == 0 || ls[g].rightPtr == 0) ? 1 : 0;
  __syncthreads();
  for (int round = 0; round <= count; round++) {
    if (threadIdx.x == 0) pending = 0;
    __syncthreads();
    // who can run: judged on the states left by the previous round
    for (int g = 1 + threadIdx.x; g <= count; g += blockDim.x) {
      if (state[g] != 0) continue;
      const int l = ls[g].leftPtr, h = ls[g].rightPtr;
      const bool ready = !(l != 0 && l < g && state[l] != 1) && !(h < g && ls[h].rightPtr != 0 && state[h] != 1);
      if (ready) state[g] = 2; else pending = 1;
    }
    __syncthreads();
    for (int g = 1 + threadIdx.x; g <= count; g += blockDim.x) {
      if (state[g] != 2) continue;
      const int h = ls[g].rightPtr;
      const float v0 = ls[g].x0, v1 = ls[g].y0, v2 = ls[g].x1, v3 = ls[g].y1;
      const float u0 = ls[h].x0, u1 = ls[h].y0, u2 = ls[h].x1, u3 = ls[h].y1;
      const float d = __fsub_rn(__fmul_rn(__fsub_rn(v2, v0), __fsub_rn(u3, u1)), __fmul_rn(__fsub_rn(v3, v1), __fsub_rn(u2, u0)));
      const float mx = __fmul_rn(__fadd_rn(v2, u0), 0.5f), my = __fmul_rn(__fadd_rn(v3, u1), 0.5f);
      float rx = mx, ry = my;
      if (!((double)fabsf(d) < 1e-6)) {
        const float n = __fsub_rn(__fmul_rn(__fsub_rn(v1, u1), __fsub_rn(u2, u0)), __fmul_rn(__fsub_rn(v0, u0), __fsub_rn(u3, u1)));
        const float q = __fdiv_rn(n, d);
        const float wx = __fadd_rn(v0, __fmul_rn(q, __fsub_rn(v2, v0))), wy = __fadd_rn(v1, __fmul_rn(q, __fsub_rn(v3, v1)));
        if (!(rd_hypot(__fsub_rn(wx, v2), __fsub_rn(wy, v3)) > 10.0f && rd_hypot(__fsub_rn(wx, u0), __fsub_rn(wy, u1)) > 10.0f)) { rx = wx; ry = wy; }
      }
      ls[g].x1 = rx; ls[g].y1 = ry;
      ls[h].x0 = rx; ls[h].y0 = ry;
      state[g] = 1;
    }
    __syncthreads();
    if (pending == 0) break;
    __syncthreads();
  }
}

// mkpl (oclpolyline.c:186-216) + refine (oclpolyline.c:299-306) for one frame per thread-block cluster.
#define PL_SYNC cluster.sync()
__global__ void __cluster_dims__(PL_CLUSTER, 1, 1) __launch_bounds__(PL_THREADS)
kp_polyline_list(LS_t *gp, int lsListSize, int *aux, int *winner, int cap, const int *table, int *flags, int *dist,
                 const int *numberin, int *labelinout, const int *plist, LSX_t *lsx, float2 *vtx, float minerror, int iw, size_t fs) {
  cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
  rd_batch_off((size_t)(blockIdx.x / PL_CLUSTER) * fs, gp, aux, winner, table, flags, dist, numberin, labelinout, plist, lsx, vtx);
  const bool lead = cluster.block_rank() == 0;
  const int N = 16;
  d_mkpl_init(gp, lsListSize, aux, winner, cap, table, flags, N);
  PL_SYNC;
  d_mkpl_pass0a(gp, lsListSize, aux, cap, numberin, labelinout, plist, iw);
  PL_SYNC;
  d_mkpl_pass0b(gp, lsListSize, aux, cap, numberin, labelinout, plist, iw);
  PL_SYNC;
  d_mkpl_pass0c(gp, aux, cap, iw);
  PL_SYNC;
  for (int it = 1; it < N; it++) {
    if (*(volatile int *)(flags + it - 1) == 0) break;   // nothing moved in the previous round: every later round is a no-op too
    d_mkpl_pass1(gp, lsListSize, dist, labelinout, plist, flags, it, iw);
    PL_SYNC;
    d_mkpl_pass2a(gp, lsListSize, winner, dist, labelinout, plist, flags, it, iw);
    PL_SYNC;
    if (lead) d_mkpl_pass2b(gp, lsListSize, winner, numberin, flags, it, minerror, iw);
    PL_SYNC;
    d_mkpl_pass3(gp, lsListSize, numberin, labelinout, plist, flags, it, iw);
    PL_SYNC;
  }
  d_refine_pass0(lsx, gp);
  PL_SYNC;
  d_refine_pass1(lsx, gp, labelinout, plist, iw);
  PL_SYNC;
  d_refine_pass2(lsx, gp);
  PL_SYNC;
  if (lead) d_refine_pass3(gp, (int *)vtx);
}

// ---------------------------------------------------------------------------- the schedule (oclpolyline.c:218-309)
static const dim3 PB(32, getenv("RD_BY") ? atoi(getenv("RD_BY")) : 8);
#define G2 rd_grid2d(iw, ih, PB)

void rd_polyline_run(LS_t *lsList, int lsListSize, int *lsIdOut, const int *in, int *tmpBig, int *tmp0, int *tmp1, int *tmp2, int *tmp3,
                     int *tmp4, int *tmp5, float minerror, int sizeThre, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  const int n = iw * ih;
  const int g1 = rd_cdiv(n, 256);
  // step 1 : clean strings
  RD_LAUNCH(kp_simpleJunction, rd_gz(G2, nb), PB, 0, s, lsIdOut, in, iw, ih, fs);
  RD_LAUNCH(kp_simpleConnect, rd_gz(G2, nb), PB, 0, s, tmp2, lsIdOut, iw, ih, fs);
  RD_LAUNCH(kp_stringify, rd_gz(G2, nb), PB, 0, s, tmp1, tmp2, 0, iw, ih, fs);
  RD_LAUNCH(kp_stringify, rd_gz(G2, nb), PB, 0, s, tmp2, tmp1, 1, iw, ih, fs);
  RD_LAUNCH(kp_removeBranch, rd_gz(G2, nb), PB, 0, s, tmp1, tmp2, iw, ih, fs);
  // step 2 : string id = smallest pixel index
  rd_label8x(lsIdOut, tmp1, tmp2, 0, iw, ih, nb, fs, s);
  // step 3 : closed loops lose their root pixel
  RD_LAUNCH(kp_simpleJunction, rd_gz(G2, nb), PB, 0, s, tmp2, tmp1, iw, ih, fs);
  rd_k_clear(tmp3, n, nb, fs, s);
  RD_LAUNCH(kp_countEnds, rd_gz(G2, nb), PB, 0, s, tmp3, tmp2, lsIdOut, iw, ih, fs);
  RD_LAUNCH(kp_breakLoops, rd_gz(G2, nb), PB, 0, s, tmp1, lsIdOut, tmp3, iw, ih, fs);
  // steps 4-6 : string ends by orientation-aware pointer jumping (8 hops x 4 launches)
  RD_LAUNCH(kp_findEnds0, rd_gz(G2, nb), PB, 0, s, tmp0, tmp2, tmpBig, lsIdOut, iw, ih, fs);
  RD_LAUNCH(kp_findEnds1, rd_gz(G2, nb), PB, 0, s, tmp3, tmp4, tmpBig, tmp0, tmp2, lsIdOut, 0, iw, ih, fs);
  RD_LAUNCH(kp_findEnds1, rd_gz(G2, nb), PB, 0, s, tmp0, tmp2, tmpBig, tmp3, tmp4, lsIdOut, 1, iw, ih, fs);
  RD_LAUNCH(kp_findEnds1, rd_gz(G2, nb), PB, 0, s, tmp3, tmp4, tmpBig, tmp0, tmp2, lsIdOut, 0, iw, ih, fs);
  RD_LAUNCH(kp_findEnds1, rd_gz(G2, nb), PB, 0, s, tmp0, tmp2, tmpBig, tmp3, tmp4, lsIdOut, 1, iw, ih, fs);
  RD_LAUNCH(kp_findEnds2, rd_gz(G2, nb), PB, 0, s, tmpBig, tmp4, tmp0, tmp2, lsIdOut, iw, ih, fs);
  // step 7 : distance from the start by list ranking (32 hops x 3 launches)
  RD_LAUNCH(kp_number, rd_gz(G2, nb), PB, 0, s, tmp2, tmp3, tmpBig, tmp4, iw, ih, fs);
  RD_LAUNCH(kp_number, rd_gz(G2, nb), PB, 0, s, tmpBig, tmp4, tmp2, tmp3, iw, ih, fs);
  RD_LAUNCH(kp_number, rd_gz(G2, nb), PB, 0, s, tmp2, tmp3, tmpBig, tmp4, iw, ih, fs);
  // step 8 : split touching strings (numbers differing by more than 1 are not connected)
  RD_LAUNCH(kp_plus1, rd_gy(g1, nb), 256, 0, s, tmp1, tmp2, n, fs);
  rd_labelpl(tmpBig, tmp1, tmp3, iw, ih, nb, fs, s);
  // step 9 : drop short strings
  rd_k_clear(tmp1, n, nb, fs, s);
  RD_LAUNCH(kp_calcSize, rd_gy(g1, nb), 256, 0, s, tmp1, tmpBig, n, fs);
  RD_LAUNCH(kp_filterSize, rd_gy(g1, nb), 256, 0, s, lsIdOut, tmpBig, tmp1, sizeThre, n, fs);
  // step 10 : compact ids 1..K in raster order of the root pixels (table = tmpBig[0..n], block counts behind it) and
  // gather the labelled pixels into a compact list (plist = tmp1, dead after filterSize): everything downstream only
  // touches those ~2 % of the frame.
  int *table = tmpBig, *blockCount = tmpBig + 2 * (size_t)n, *plist = tmp1;
  {
    const int nblk = rd_cdiv(n, RL_BLOCK);
    rd_k_clear(tmpBig, n + 1, nb, fs, s);
    RD_LAUNCH(kp_relabel_count, rd_gy(nblk, nb), RL_BLOCK, 0, s, blockCount, lsIdOut, iw, ih, fs);
    RD_LAUNCH(kp_scan_blocks, dim3(nb), 1024, 0, s, blockCount, nblk, table, plist, fs);
    RD_LAUNCH(kp_relabel_rank, rd_gy(nblk, nb), RL_BLOCK, 0, s, table, blockCount, lsIdOut, iw, ih, fs);
    RD_LAUNCH(kp_relabel_pass1, rd_gz(G2, nb), PB, 0, s, lsIdOut, table, plist, iw, ih, fs);
  }
  // steps 11-12 : mkpl and refine in ONE launch (kp_polyline_list: one CTA per frame walks the compact pixel list and
  // the segment list; the 60-odd launches of the reference's split loop are __syncthreads here).
  // aux (start / end pixel per string) and winner (arg-max pixel per segment) sit behind the table; the LSX mirror of
  // the list is at the start of tmpBig (the table is dead by then), shared vertices in tmp3 (the distance slots are dead by then).
  {
    const int cap = lsListSize / (int)sizeof(LS_t);
    int *aux = tmpBig + n + 8, *winner = aux + 2 * (size_t)cap;
    RD_LAUNCH(kp_polyline_list, dim3(nb * PL_CLUSTER), PL_THREADS, 0, s, lsList, lsListSize, aux, winner, cap, table, tmp4, tmp3, tmp2, lsIdOut, plist, (LSX_t *)tmpBig,
              (float2 *)tmp3, minerror, iw, fs);
  }
  (void)tmp5;
}

extern "C" {

oclpolyline_t *init_oclpolyline(cl_device_id device, cl_context) {          // oclpolyline.c:20-101
  if (rd_device_count() <= 0) exitf(-1, "rectdetect_b200: no CUDA device; there is no CPU fallback\n");
  oclpolyline_t *t = (oclpolyline_t *)calloc(1, sizeof(oclpolyline_t));
  t->magic = POLY_MAGIC;
  t->ordinal = device ? device->ordinal : 0;
  return t;
}
void dispose_oclpolyline(oclpolyline_t *thiz) {
  if (!thiz || thiz->magic != POLY_MAGIC) exitf(-1, "rectdetect_b200: bad oclpolyline_t\n");
  thiz->magic = 0;
  free(thiz);
}

cl_event oclpolyline_execute(oclpolyline_t *thiz, cl_mem lsList, int lsListSize, cl_mem lsIdOut, cl_mem in, cl_mem tmp0, cl_mem tmp1, cl_mem tmp2,
                             cl_mem tmp3, cl_mem tmp4, cl_mem tmp5, cl_mem tmp6, float minerror, int sizeThre, int iw, int ih,
                             cl_command_queue queue, const cl_event *events) {
  if (!thiz || thiz->magic != POLY_MAGIC) exitf(-1, "rectdetect_b200: bad oclpolyline_t\n");
  cudaStream_t s = rd_stream(queue);
  rd_wait_events(s, events);
  const size_t P = (size_t)iw * ih * 4;
  rd_need(lsList, (size_t)lsListSize, "oclpolyline_execute lsList");
  rd_need(lsIdOut, P, "oclpolyline_execute lsIdOut"); rd_need(in, P, "oclpolyline_execute in");
  rd_need(tmp0, 4 * P, "oclpolyline_execute tmpBig");
  cl_mem t[6] = {tmp1, tmp2, tmp3, tmp4, tmp5, tmp6};
  for (int i = 0; i < 6; i++) rd_need(t[i], P, "oclpolyline_execute tmp");
  if ((size_t)lsListSize > 4 * P) exitf(-1, "rectdetect_b200: oclpolyline_execute needs lsListSize <= iw*ih*16\n");
  rd_polyline_run(rd_ptr<LS_t>(lsList), lsListSize, rd_ptr<int>(lsIdOut), rd_ptr<int>(in), rd_ptr<int>(tmp0), rd_ptr<int>(tmp1), rd_ptr<int>(tmp2),
                  rd_ptr<int>(tmp3), rd_ptr<int>(tmp4), rd_ptr<int>(tmp5), rd_ptr<int>(tmp6), minerror, sizeThre, iw, ih, 1, 0, s);
  return rd_make_event(s, events);
}

}  // extern "C"

// =============================================================================================== the rect pipeline's polyline stage
// rd_polyline_fast computes what rd_polyline_run computes (same lsIdOut map, same list) for the device schedule of
// rd_rect.cu, where the stale frame simpleConnect leaves behind (oclpolyline.cl:91) is known to be non-zero:
//   - the five string clean-up kernels of step 1 are one shared-memory kernel on byte tiles;
//   - the foreground of the string labelling is gathered into a compact pixel list, and end finding, numbering, the
//     loop breaker, size filter and relabelling only touch listed pixels (a few % of the frame) instead of sweeping planes;
//   - ids are handed out in raster order of the root pixels by ranking the (few hundred) roots against each other.
// Bit-plane tile (rd_bits.cuh), apron 6 = junction 1 + connect 2 + stringify 1 + 1 + removeBranch 1.
#define KB2_A 6
#define KB2_R (BT_PR + 2 * KB2_A)
struct IntNonZero { __device__ __forceinline__ bool operator()(int v) const { return v != 0; } };
__global__ void __launch_bounds__(256) kb_strings2(uint8_t *out, int *copyOut, int *list, const int *strong, int ring, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, copyOut, list, strong);
  __shared__ bt_plane pa[KB2_R], pz[KB2_R], pt[KB2_R];
  const int bx0 = blockIdx.x * (32 * BT_PW), by0 = blockIdx.y * BT_PR, gy0 = by0 - KB2_A;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) list[0] = 0;
  {
    // the copy for the next frame's strength accumulator is taken from the payload rows only (every pixel exactly once)
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    if (copyOut)
      for (int pr = wy; pr < BT_PR; pr += 8) {
        const int gy = by0 + pr;
        if (gy >= ih) break;
        for (int k = 0; k < BT_PW; k++) { const int gx = bx0 + 32 * k + lane; if (gx < iw) copyOut[(size_t)gy * iw + gx] = strong[(size_t)gy * iw + gx]; }
      }
  }
  bt_build(pa, strong, KB2_R, bx0, gy0, iw, ih, IntNonZero());
  __syncthreads();
  // simpleJunction (oclpolyline.cl:66): pz = value != 0, pt = value == 2
  BT_TASKS(KB2_R) {
    BT_RC;
    uint32_t nz = 0, eq2 = 0;
    if (r >= 1 && r < KB2_R - 1) {
      const int gy = gy0 + r, gx0 = bx0 + 32 * (c - 1);
      const BtNb nb = bt_neighbours<false>(pa, r, c);
      nz = nb.centre & nb.any & (bt_rowok(gy, ih, 1) ? bt_cols(gx0, iw, 1) : 0u);
      eq2 = nz & ~nb.ge2;
    }
    pz[r][c] = nz; pt[r][c] = eq2;
  }
  __syncthreads();
  // simpleConnect (oclpolyline.cl:89): bridges one-pixel gaps between two end pixels; the 2-px frame keeps what was there (`ring`)
  BT_TASKS(KB2_R) {
    BT_RC;
    uint32_t res = 0;
    if (r >= 3 && r < KB2_R - 3) {
      const int gy = gy0 + r, gx0 = bx0 + 32 * (c - 1);
      const uint32_t inimg = (gy >= 0 && gy < ih) ? bt_cols(gx0, iw, 0) : 0u, in2 = bt_rowok(gy, ih, 2) ? bt_cols(gx0, iw, 2) : 0u;
      const Bt3 z2n = bt_load3(pz, r - 2, c), zm = bt_load3(pz, r, c), zs = bt_load3(pz, r + 1, c), z2s = bt_load3(pz, r + 2, c);
      const Bt3 tn = bt_load3(pt, r - 1, c), tm = bt_load3(pt, r, c), ts = bt_load3(pt, r + 1, c);
      // Z(dx, dy) / T(dx, dy): the plane at (x + dx, y + dy)
      const uint32_t p1 = bt_w(zm, 2) & bt_w(tm, 1) & bt_e(tm, 1) & bt_e(zm, 2);
      const uint32_t p2 = z2n.c & tn.c & ts.c & z2s.c;
      const uint32_t p3 = bt_w(z2n, 2) & bt_w(tn, 1) & bt_e(ts, 1) & bt_e(z2s, 2);
      const uint32_t p4 = bt_e(z2n, 2) & bt_e(tn, 1) & bt_w(ts, 1) & bt_w(z2s, 2);
      const uint32_t p5 = bt_e(zm, 2) & bt_e(tm, 1) & bt_w(ts, 1) & bt_w(zs, 2);
      const uint32_t p6 = bt_w(zm, 2) & bt_w(tm, 1) & bt_e(ts, 1) & bt_e(zs, 2);
      const uint32_t p7 = bt_e(z2n, 1) & bt_e(tn, 1) & ts.c & z2s.c;
      const uint32_t p8 = bt_w(z2n, 1) & bt_w(tn, 1) & ts.c & z2s.c;
      res = (in2 & (zm.c | p1 | p2 | p3 | p4 | p5 | p6 | p7 | p8)) | (ring ? (inimg & ~in2) : 0u);
    }
    pa[r][c] = res;
  }
  __syncthreads();
  // stringify 0, 1 (oclpolyline.cl:112)
  BT_TASKS(KB2_R) {
    BT_RC;
    pz[r][c] = (r >= 4 && r < KB2_R - 4) ? bt_stringify(pa, r, c, bx0 + 32 * (c - 1), gy0 + r, iw, ih, 0) : 0u;
  }
  __syncthreads();
  BT_TASKS(KB2_R) {
    BT_RC;
    pa[r][c] = (r >= 5 && r < KB2_R - 5) ? bt_stringify(pz, r, c, bx0 + 32 * (c - 1), gy0 + r, iw, ih, 1) : 0u;
  }
  __syncthreads();
  // removeBranch (oclpolyline.cl:126): only pixels with at most two neighbours stay
  BT_TASKS(KB2_R) {
    BT_RC;
    uint32_t res = 0;
    if (r >= KB2_A && r < KB2_R - KB2_A) {
      const int gy = gy0 + r, gx0 = bx0 + 32 * (c - 1);
      const BtNb nb = bt_neighbours<true>(pa, r, c);
      res = nb.centre & ~nb.ge3 & (bt_rowok(gy, ih, 1) ? bt_cols(gx0, iw, 1) : 0u);
    }
    pz[r][c] = res;
  }
  __syncthreads();
  bt_store_bytes(out, pz, KB2_A, bx0, by0, iw, ih);
}

// ---- list kernels: grid-stride over the compact list of string pixels (frame = blockIdx.y) ----
#define SL_LOOP const int scount_ = list[0]; for (int k_ = blockIdx.x * blockDim.x + threadIdx.x; k_ < scount_; k_ += gridDim.x * blockDim.x)
#define SL_P const int p0 = list[k_ + 1]; const int x = p0 % iw, y = p0 / iw; (void)x; (void)y

// loop breaker, part 1 of 3 (oclpolyline.cl:149-167): roots start out with "no end seen"
__global__ void kl_ends_reset(int *nEnds, const int *label, const int *list, int iw, size_t fs) {
  rd_batch_y(fs, nEnds, label, list);
  SL_LOOP { SL_P; if (label[p0] == p0) nEnds[p0] = 0; }
}
// part 2: a string pixel with exactly one neighbour (simpleJunction value 2) marks its string as open
__global__ void kl_ends_mark(int *nEnds, const uint8_t *str, const int *label, const int *list, int iw, size_t fs) {
  rd_batch_y(fs, nEnds, str, label, list);
  SL_LOOP {
    SL_P;
    int c = 1;
#pragma unroll
    for (int i = 0; i < 8; i++) c += str[p0 + RD_RX[i] + RD_RY[i] * iw] != 0;
    if (c == 2) nEnds[label[p0]] = 1;
  }
}
// part 3: closed loops lose their root pixel
__global__ void kl_break_loops(uint8_t *str, int *label, const int *nEnds, const int *list, int iw, size_t fs) {
  rd_batch_y(fs, str, label, nEnds, list);
  SL_LOOP { SL_P; if (label[p0] == p0 && nEnds[p0] == 0) { str[p0] = 0; label[p0] = -1; } }
}
__global__ void kl_findEnds0(int *nextout, int *prevout, int *flagout, const int *labelin, const int *list, int iw, size_t fs) {
  rd_batch_y(fs, nextout, prevout, flagout, labelin, list);
  SL_LOOP {
    SL_P;
    if (labelin[p0] == -1) continue;
    int npx, npy, a, b, flag = 0;
    getnp(labelin, p0, iw, npx, npy);
    if (npx != p0) { getnp(labelin, npx, iw, a, b); if (a == p0) flag |= 1; }
    if (npy != p0) { getnp(labelin, npy, iw, a, b); if (b == p0) flag |= 2; }
    nextout[p0] = npx; prevout[p0] = npy; flagout[p0] = flag;
  }
}
__global__ void kl_findEnds1(int *nextout, int *prevout, int *flaginout, const int *nextin, const int *previn, const int *labelin, const int *list, int page, int iw, size_t fs) {
  rd_batch_y(fs, nextout, prevout, flaginout, nextin, previn, labelin, list);
  SL_LOOP {
    SL_P;
    if (labelin[p0] == -1) continue;
    const volatile int *fl = flaginout;
    const int f0 = fl[p0];
    bool revn = page == 0 ? ((f0 & 1) != 0) : ((f0 & 4) != 0);
    bool revp = page == 0 ? ((f0 & 2) != 0) : ((f0 & 8) != 0);
    int nn = nextin[p0], pp = previn[p0];
    for (int i = 0; i < 8; i++) {
      const int nn2 = revn ? previn[nn] : nextin[nn];
      const int pp2 = revp ? nextin[pp] : previn[pp];
      int nflag = fl[nn], pflag = fl[pp];
      if (page != 0) { nflag >>= 2; pflag >>= 2; }
      revn = revn ? ((nflag & 2) == 0) : ((nflag & 1) != 0);
      revp = revp ? ((pflag & 1) == 0) : ((pflag & 2) != 0);
      nn = nn2; pp = pp2;
    }
    int f = f0;
    if (page == 0) { f &= 3; f |= revn ? 4 : 0; f |= revp ? 8 : 0; }
    else { f &= (3 << 2); f |= revn ? 1 : 0; f |= revp ? 2 : 0; }
    flaginout[p0] = f;
    nextout[p0] = nn; prevout[p0] = pp;
  }
}
__global__ void kl_findEnds2(int *numout, int *linkout, const int *nextin, const int *previn, const int *labelin, const int *list, int iw, size_t fs) {
  rd_batch_y(fs, numout, linkout, nextin, previn, labelin, list);
  SL_LOOP {
    SL_P;
    int num = 0, link = -1;
    if (labelin[p0] != -1) {
      int npx, npy;
      getnp(labelin, p0, iw, npx, npy);
      link = nextin[p0] < previn[p0] ? npx : npy;
      num = link == p0 ? 0 : 1;
    }
    numout[p0] = num; linkout[p0] = link;
  }
}
// pl1 (optional): number + 1 where the number is non-zero, for the labelpl step (plane cleared by the caller)
__global__ void kl_number(int *numout, int *linkout, int *pl1, const int *numin, const int *linkin, const int *list, int iw, int npix, size_t fs) {
  rd_batch_y(fs, numout, linkout, pl1, numin, linkin, list);
  SL_LOOP {
    SL_P;
    int no = 0, lo = -1;
    const int l0 = linkin[p0];
    if (l0 == -1) { no = numin[p0]; lo = -1; }
    else {
      int n = numin[p0], l = l0;
      bool bail = false;
      for (int i = 0; i < 32; i++) {
        if (!(0 < l && l < npix)) { bail = true; break; }
        n += numin[l];
        l = linkin[l];
      }
      if (!bail) { no = n; lo = l; }
    }
    numout[p0] = no; linkout[p0] = lo;
    if (pl1 && no != 0) pl1[p0] = no + 1;
  }
}
// calcSize / filterSize (oclpolyline.cl:357-378) on the list; roots are also gathered (roots[0] = count) for the ranking
__global__ void kl_size_reset(int *size, int *roots, const int *lab, const int *list, int iw, size_t fs) {
  rd_batch_y(fs, size, roots, lab, list);
  if (blockIdx.x == 0 && threadIdx.x == 0) roots[0] = 0;
  SL_LOOP { SL_P; if (lab[p0] == p0) size[p0] = 0; }
}
__global__ void kl_size_count(int *size, const int *lab, const int *list, int iw, size_t fs) {
  rd_batch_y(fs, size, lab, list);
  SL_LOOP { SL_P; const int b = lab[p0]; if (b != 0) atomicAdd(size + b, 1); }
}
__global__ void kl_roots(int *roots, const int *size, const int *lab, const int *list, int sizeThre, int iw, size_t fs) {
  rd_batch_y(fs, roots, size, lab, list);
  SL_LOOP { SL_P; if (lab[p0] == p0 && size[p0] > sizeThre) roots[1 + atomicAdd(roots, 1)] = p0; }
}
// ids 1..K in raster order of the surviving roots (relabel_pass0, oclpolyline.cl:380, canonical order Q4); table[0] = K
__global__ void kl_rank(int *table, const int *roots, size_t fs) {
  rd_batch_y(fs, table, roots);
  const int K = roots[0];
  if (blockIdx.x == 0 && threadIdx.x == 0) table[0] = K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K; i += gridDim.x * blockDim.x) {
    const int r = roots[1 + i];
    int rank = 1;
    for (int j = 0; j < K; j++) rank += roots[1 + j] < r;
    table[r + 1] = rank;
  }
}
// filterSize + relabel_pass1 (oclpolyline.cl:367, 400): lsIdOut was cleared by the caller
__global__ void kl_relabel(int *lsIdOut, const int *table, const int *size, const int *lab, const int *list, int sizeThre, int iw, size_t fs) {
  rd_batch_y(fs, lsIdOut, table, size, lab, list);
  SL_LOOP { SL_P; const int b = lab[p0]; if (b != 0 && size[b] > sizeThre) lsIdOut[p0] = table[b + 1]; }
}

void rd_label8x_u8_list(int *label, const uint8_t *pix, void *scratch, int *list, int bgc, int iw, int ih, int nb, size_t fs, cudaStream_t s);

// in: strong-edge bitmap; copyOut (optional): receives a copy of `in`; t0..t5: scratch planes; tmpBig: 4 planes
void rd_polyline_fast(LS_t *lsList, int lsListSize, int *lsIdOut, const int *in, int *copyOut, int *tmpBig, int *t0, int *t1, int *t2, int *t3, int *t4, int *t5,
                      float minerror, int sizeThre, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  const int n = iw * ih;
  uint8_t *str = (uint8_t *)t0;
  int *roots = t0 + (n + 3) / 4 + 4;                       // behind the string bytes
  int *list = t2;
  int *nextA = t3, *prevA = t4, *nextB = tmpBig, *prevB = tmpBig + n, *flags1 = tmpBig + 2 * (size_t)n;
  const dim3 lg(LIST_BLOCKS, nb);
  // step 1 : string clean-up, one kernel (also zeroes the list counter and copies the bitmap for the next frame)
  RD_LAUNCH(kb_strings2, dim3(rd_cdiv(iw, 32 * BT_PW), rd_cdiv(ih, BT_PR), nb), 256, 0, s, str, copyOut, list, in, 1, iw, ih, fs);
  // step 2 : string id = smallest pixel index; foreground gathered into `list`
  rd_label8x_u8_list(lsIdOut, str, t1, list, 0, iw, ih, nb, fs, s);
  // step 3 : closed loops lose their root pixel
  RD_LAUNCH(kl_ends_reset, lg, 256, 0, s, t5, lsIdOut, list, iw, fs);
  RD_LAUNCH(kl_ends_mark, lg, 256, 0, s, t5, str, lsIdOut, list, iw, fs);
  RD_LAUNCH(kl_break_loops, lg, 256, 0, s, str, lsIdOut, t5, list, iw, fs);
  // steps 4-6 : string ends
  RD_LAUNCH(kl_findEnds0, lg, 256, 0, s, nextA, prevA, flags1, lsIdOut, list, iw, fs);
  RD_LAUNCH(kl_findEnds1, lg, 256, 0, s, nextB, prevB, flags1, nextA, prevA, lsIdOut, list, 0, iw, fs);
  RD_LAUNCH(kl_findEnds1, lg, 256, 0, s, nextA, prevA, flags1, nextB, prevB, lsIdOut, list, 1, iw, fs);
  RD_LAUNCH(kl_findEnds1, lg, 256, 0, s, nextB, prevB, flags1, nextA, prevA, lsIdOut, list, 0, iw, fs);
  RD_LAUNCH(kl_findEnds1, lg, 256, 0, s, nextA, prevA, flags1, nextB, prevB, lsIdOut, list, 1, iw, fs);
  int *numA = tmpBig, *linkA = tmpBig + n;                 // (nextB / prevB are dead)
  RD_LAUNCH(kl_findEnds2, lg, 256, 0, s, numA, linkA, nextA, prevA, lsIdOut, list, iw, fs);
  // step 7 : distance from the start; the last round also writes number + 1 into the cleared plane the next labelling reads
  int *numB = t3, *linkB = t4, *pl1 = t5;
  RD_LAUNCH(kl_number, lg, 256, 0, s, numB, linkB, (int *)NULL, numA, linkA, list, iw, n, fs);
  RD_LAUNCH(kl_number, lg, 256, 0, s, numA, linkA, (int *)NULL, numB, linkB, list, iw, n, fs);
  rd_k_clear(pl1, n, nb, fs, s);
  RD_LAUNCH(kl_number, lg, 256, 0, s, numB, linkB, pl1, numA, linkA, list, iw, n, fs);
  int *number = numB;                                      // t3, survives to the end
  // step 8 : split touching strings
  int *lab3 = tmpBig + 3 * (size_t)n;
  rd_labelpl(lab3, pl1, t1, iw, ih, nb, fs, s);
  // steps 9-10 : drop short strings, ids 1..K in raster order of the roots
  int *size = t4, *table = tmpBig;
  RD_LAUNCH(kl_size_reset, lg, 256, 0, s, size, roots, lab3, list, iw, fs);
  RD_LAUNCH(kl_size_count, lg, 256, 0, s, size, lab3, list, iw, fs);
  RD_LAUNCH(kl_roots, lg, 256, 0, s, roots, size, lab3, list, sizeThre, iw, fs);
  RD_LAUNCH(kl_rank, dim3(4, nb), 256, 0, s, table, roots, fs);
  rd_k_clear(lsIdOut, n, nb, fs, s);
  RD_LAUNCH(kl_relabel, lg, 256, 0, s, lsIdOut, table, size, lab3, list, sizeThre, iw, fs);
  // steps 11-12 : mkpl + refine on the same list (pixels without an id are skipped there)
  {
    const int cap = lsListSize / (int)sizeof(LS_t);
    int *aux = tmpBig + n + 8, *winner = aux + 2 * (size_t)cap;
    RD_LAUNCH(kp_polyline_list, dim3(nb * PL_CLUSTER), PL_THREADS, 0, s, lsList, lsListSize, aux, winner, cap, table, t5, t4, number, lsIdOut, list, (LSX_t *)tmpBig, (float2 *)t4,
              minerror, iw, fs);
  }
}
