// rd_rect.cu - Stage B and D of the rectangle detector (oclrect.cl) and the oclrect_t pipeline object
// (oclrect.h:17-23, oclrect.c:41-381, 1230-1278) on sm_100a.
//
// The device schedule (gpu_task) replays genGPUTask (oclrect.c:235-381) step for step on one CUDA stream with the
// reference's buffer plan (6 buf + 6 tmp + 2 iobuf planes, 2 ioBig), so the two cross-frame carry-overs of the
// reference (strengths accumulate on top of the previous strong-edge mask; Q2: region sizes accumulate on
// top of the junction map) and the stale frame that oclpolyline.cl's simpleConnect leaves in its output are
// reproduced by construction.  What differs from the reference, on purpose:
//   - the bounded label propagation loops are exact connected components (rd_ccl.cu);
//   - nothing big is read back: executeCPUTask (oclrect.c:1049-1226) runs on the device too (rd_gtail.cu), in IEEE double in
//     the reference's order of operations, so a frame costs one small D2H copy - its rect_t list - instead of the
//     reference's 9 planes (33 MB at 1280x720), and no host core touches it.
#include <chrono>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include "rd_common.cuh"
#include "rd_stageA.cuh"

typedef linesegment_t LS_t;

void rd_label8x(int *label, const int *pix, void *scratch, int bgc, int iw, int ih, int nb, size_t fs, cudaStream_t s);
void rd_labelMerge(int *out, int *work, const uint32_t *pix, const int *mask, const int *edge, void *scratch, int *flags, int iw, int ih, int nb, size_t fs, cudaStream_t s);
extern "C" int rd_get_merge_replay(void);
void rd_k_clear(int *out, int nints, int nb, size_t fs, cudaStream_t s);
void rd_k_copy(int *out, const int *in, int nints, int nb, size_t fs, cudaStream_t s);
void rd_k_iirblur(float *obuf, const float *ibuf, float *tmp0, float *tmp1, int r, int iw, int ih, int nb, size_t fs, cudaStream_t s);
void rd_polyline_run(LS_t *lsList, int lsListSize, int *lsIdOut, const int *in, int *tmpBig, int *tmp0, int *tmp1, int *tmp2, int *tmp3,
                     int *tmp4, int *tmp5, float minerror, int sizeThre, int iw, int ih, int nb, size_t fs, cudaStream_t s);
void rd_bgr2plab_run(uint32_t *out, const uint8_t *in, size_t in_fs, int iw, int ih, int ws, int nb, size_t fs, cudaStream_t s);
void rd_iirblur3_run(float *outL, float *outA, float *outB, uint32_t *outPlab, const uint32_t *plab, float *sb, float *sf, size_t pp, int r, int iw, int ih,
                     int nb, size_t fs, cudaStream_t s);
void rd_blblur_run(uint32_t *dst, uint32_t *pong, const uint32_t *src, const int8_t *edge, uint8_t *ext, int iters, int iw, int ih, int nb, size_t fs, cudaStream_t s);
void rd_calcSize_run(int *out, const int *label, int iw, int ih, int nb, size_t fs, cudaStream_t s);
void rd_edge_thin_run(float *thin, const float *blurL, const uint32_t *blurP, int iw, int ih, int nb, size_t fs, cudaStream_t s);
void rd_strings1_run(uint8_t *out, const float *thin, int iw, int ih, int nb, size_t fs, cudaStream_t s);
void rd_filter_masks_run(int8_t *weak, int *strong, const int *label, const int *str, int iw, int ih, int nb, size_t fs, cudaStream_t s);
void rd_quant_despeckle_run(uint32_t *out, const uint32_t *in, const float *thin, int iw, int ih, int nb, size_t fs, cudaStream_t s);
void rd_quant_tables_init();
void rd_junction_mask_run(uint8_t *mask, int *junc, const int *strong, int iw, int ih, int nb, size_t fs, cudaStream_t s);
void rd_despeckle2_run(int *dst, const int *label, const int *size, int *list, int *recL, int *recS, int *rowcnt, int2 *rowbuf, int thre, int iw, int ih,
                       int nb, size_t fs, cudaStream_t s);
void rd_markBoundary_run(int *out, const int *in, int iw, int ih, int nb, size_t fs, cudaStream_t s);
void rd_label8x_u8(int *label, const uint8_t *pix, void *scratch, int bgc, int iw, int ih, int nb, size_t fs, cudaStream_t s);
void rd_labelMerge_u8(int *out, int *work, const uint32_t *pix, const uint8_t *mask, const int *edge, void *scratch, int *flags, int iw, int ih, int nb, size_t fs, cudaStream_t s);
void rd_polyline_fast(LS_t *lsList, int lsListSize, int *lsIdOut, const int *in, int *copyOut, int *tmpBig, int *t0, int *t1, int *t2, int *t3, int *t4, int *t5,
                      float minerror, int sizeThre, int iw, int ih, int nb, size_t fs, cudaStream_t s);
void rd_gtail_run(unsigned char *blob, size_t blobBytes, const linesegment_t *ls, const int *segid, const int *votes, int *table, unsigned char *scratch,
                  size_t scratchBytes, int iw, int ih, double tanAOV, int phases, int nb, size_t fs, cudaStream_t s);

#define XY2D const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y; if (x >= iw || y >= ih) return; const int p0 = y * iw + x
static const dim3 RB(32, getenv("RD_BY") ? atoi(getenv("RD_BY")) : 8);
#define G2 rd_grid2d(iw, ih, RB)

// ---------------------------------------------------------------------------- oclrect.cl:74-135
__global__ void kr_simpleJunction(int *out, const int *in, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, in);
  XY2D;
  int r = 0;
  if (!(x <= 0 || y <= 0 || x >= iw - 1 || y >= ih - 1) && in[p0] > 0) {
    int count = 1;
#pragma unroll
    for (int i = 0; i < 8; i++) if (in[p0 + RD_RX[i] + RD_RY[i] * iw] > 0) count++;
    r = count == 1 ? 0 : count;
  }
  out[p0] = r;
}
__global__ void kr_simpleConnect(int *out, const int *in, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, in);
  XY2D;
  int r = 0;
  if (!(x <= 1 || y <= 1 || x >= iw - 2 || y >= ih - 2)) {
    r = in[p0] != 0 ? 1 : 0;
    if (!r) {
      const int w = in[p0 - 1], e = in[p0 + 1], n = in[p0 - iw], s = in[p0 + iw];
      const int nw = in[p0 - iw - 1], ne = in[p0 - iw + 1], sw = in[p0 + iw - 1], se = in[p0 + iw + 1];
      if (w == 2 && e != 0) r = 1;
      if (w != 0 && e == 2) r = 1;
      if (n == 2 && s != 0) r = 1;
      if (n != 0 && s == 2) r = 1;
      if (nw == 2 && se == 2) r = 1;
      if (ne == 2 && sw == 2) r = 1;
      if (e == 2 && sw == 2) r = 1;
      if (w == 2 && se == 2) r = 1;
      if (ne == 2 && s == 2) r = 1;
      if (nw == 2 && s == 2) r = 1;
    }
  }
  out[p0] = r;
}
__global__ void kr_stringify(int *out, const int *in, int mod2, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, in);
  XY2D;
  int r = in[p0];
  if (!(x <= 0 || y <= 0 || x >= iw - 1 || y >= ih - 1) && ((x + y) & 1) == mod2) {
    const bool n = in[p0 - iw] != 0, s = in[p0 + iw] != 0, w = in[p0 - 1] != 0, e = in[p0 + 1] != 0;
    if ((n || s) && (w || e)) r = 0;
  }
  out[p0] = r;
}

// ---------------------------------------------------------------------------- oclrect.cl:137-153 (same code as oclimgutil.cl:641-657)
__global__ void kr_calcStrength(int *out, const float *edge, const int *label, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, edge, label);
  XY2D;
  if (x <= 0 || y <= 0 || x >= iw - 1 || y >= ih - 1) return;
  const int l = label[p0];
  if (l <= 0) return;
  const float e = edge[p0];
  const int v = (int)__fmul_rn(__fmul_rn(e, e), 10000.0f);
  if (v != 0) atomicAdd(out + l, v);
}
__global__ void kr_filterStrength(int *labelinout, const int *str, int thre, int iw, int ih, size_t fs) {
  rd_batch_z(fs, labelinout, str);
  XY2D;
  if (x <= 0 || y <= 0 || x >= iw - 1 || y >= ih - 1) return;
  const int l = labelinout[p0];
  if (l <= 0 || str[l] < thre) labelinout[p0] = -1;
}
__global__ void kr_threshold_i_i(int *out, const int *in, int vlow, int thr, int vhigh, int n, size_t fs) {
  rd_batch_y(fs, out, in);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] > thr ? vhigh : vlow;
}
__global__ void kr_threshold_cast(float *outf, int *outi, const float *in, int n, size_t fs) {
  rd_batch_y(fs, outf, outi, in);     // threshold_f_f(0,0,1) + cast_i_f(1.0), oclrect.c:262-263
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float t = in[i] > 0.0f ? 1.0f : 0.0f;
  outf[i] = t;
  outi[i] = (int)t;
}
__global__ void kr_threshold_cast_c(int *outi, int8_t *outc, const int *in, int n, size_t fs) {
  rd_batch_y(fs, outi, outc, in);    // threshold_i_i(0,0,1) + cast_c_i, oclrect.c:282-284
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = in[i] > 0 ? 1 : 0;
  outi[i] = t;
  outc[i] = (int8_t)t;
}

// ---------------------------------------------------------------------------- oclrect.cl:155-205 : edge-stopped box blur
#define BLBLURSIZE 4
__device__ __forceinline__ uint32_t packlabbl(int l, int a, int b) {
  uint32_t ret = (uint32_t)min(max(b, 0), 1023);
  ret = (ret << 10) | (uint32_t)min(max(a, 0), 1023);
  ret = (ret << 12) | (uint32_t)min(max(l, 0), 4095);
  return ret;
}
// DIR 0: along x (blblur0), DIR 1: along y (blblur1).  `step` is the stride of the walk, `side` the stride of the
// perpendicular neighbour the second stop rule looks at (row below for blblur0, column to the right for blblur1).
template <int DIR>
__global__ void kr_blblur(uint32_t *out, const int8_t *edge, const uint32_t *in, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, edge, in);
  XY2D;
  const int pos = DIR == 0 ? x : y, len = DIR == 0 ? iw : ih;
  const int step = DIR == 0 ? 1 : iw, side = DIR == 0 ? iw : 1;
  const bool hasSide = DIR == 0 ? (y < ih - 1) : (x < iw - 1);
  int wsum = 0, c0 = 0, c1 = 0, c2 = 0;
  const int oe = edge[p0] != 0;
  for (int d = 0; d >= -BLBLURSIZE; d--) {
    if (pos + d < 0) break;
    const int q = p0 + d * step;
    if (pos + d > 0 && edge[q] != 0 && edge[q - step] == 0) break;
    if (pos + d > 0 && hasSide && edge[q] == 0 && edge[q - step] != 0 && edge[q + side] != 0) break;
    wsum++;
    const uint32_t v = in[q];
    c0 += v & 4095; c1 += (v >> 12) & 1023; c2 += (v >> 22) & 1023;
  }
  for (int d = 0; d <= BLBLURSIZE; d++) {
    if (pos + d > len - 1) break;
    const int q = p0 + d * step;
    if (pos + d < len - 1 && edge[q] == 0 && edge[q + step] != 0) break;
    if (oe && edge[q] == 0) break;
    wsum++;
    const uint32_t v = in[q];
    c0 += v & 4095; c1 += (v >> 12) & 1023; c2 += (v >> 22) & 1023;
  }
  out[p0] = wsum == 0 ? in[p0] : packlabbl(c0 / wsum, c1 / wsum, c2 / wsum);
}

// ---------------------------------------------------------------------------- oclrect.cl:207-244
__global__ void kr_quantize(uint32_t *out, const uint32_t *in, int n0, int n1, int n2, int n, size_t fs) {
  rd_batch_y(fs, out, in);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float l, a, b;
  rd_unpacklab(in[i], l, a, b);
  out[i] = rd_packlab(__fdiv_rn(roundf(__fmul_rn(l, (float)n0)), (float)n0), __fdiv_rn(roundf(__fmul_rn(a, (float)n1)), (float)n1),
                      __fdiv_rn(roundf(__fmul_rn(b, (float)n2)), (float)n2));
}
__global__ void kr_despeckle(uint32_t *out, const uint32_t *in, const float *edge, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, in, edge);
  XY2D;
  uint32_t r = in[p0];
  if (!(edge[p0] < 1e-6f)) {
    float dist = 1e+10f, l0, a0, b0;
    rd_unpacklab(r, l0, a0, b0);
    for (int yy = -1; yy <= 1; yy++)
      for (int xx = -1; xx <= 1; xx++)
        if (0 <= x + xx && x + xx < iw && 0 <= y + yy && y + yy < ih) {
          const int p1 = (y + yy) * iw + x + xx;
          if (edge[p1] >= 1e-6f) continue;
          float l1, a1, b1;
          const uint32_t v = in[p1];
          rd_unpacklab(v, l1, a1, b1);
          const float d = rd_distance3(__fsub_rn(l1, l0), __fsub_rn(a1, a0), __fsub_rn(b1, b0));
          if (d < dist) { r = v; dist = d; }
        }
  }
  out[p0] = r;
}

// ---------------------------------------------------------------------------- oclrect.cl:246-287 : scatters of constants (order independent)
__global__ void kr_mkMergeMask0(int *out, const int *junctionIn, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, junctionIn);
  XY2D;
  if (junctionIn[p0] == 0) return;
  for (int yy = max(y - 6, 0); yy <= min(y + 6, ih - 1); yy++)
    for (int xx = max(x - 6, 0); xx <= min(x + 6, iw - 1); xx++) {
      const int dsqu = (yy - y) * (yy - y) + (xx - x) * (xx - x);
      if (16 <= dsqu && dsqu < 36) out[yy * iw + xx] = 1;
    }
}
__global__ void kr_mkMergeMask1(int *inout, const int *junctionIn, int iw, int ih, size_t fs) {
  rd_batch_z(fs, inout, junctionIn);
  XY2D;
  const int j = junctionIn[p0];
  if (j == 0) return;
  const int r = j == 2 ? 8 : 4, lim = j == 2 ? 64 : 16;
  for (int yy = max(y - r, 0); yy <= min(y + r, ih - 1); yy++)
    for (int xx = max(x - r, 0); xx <= min(x + r, iw - 1); xx++) {
      const int dsqu = (yy - y) * (yy - y) + (xx - x) * (xx - x);
      if (dsqu < lim) inout[yy * iw + xx] = 0;
    }
}

// ---------------------------------------------------------------------------- oclrect.cl:336-390
__global__ void kr_calcSize(int *out, const int *label, int n, size_t fs) {
  rd_batch_y(fs, out, label);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int l = label[i];
  if (l != -1) atomicAdd(out + l, 1);
}
__global__ void kr_markBoundary(int *out, const int *in, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, in);
  XY2D;
  int r = -1;
  if (!(x <= 1 || y <= 1 || x >= iw - 2 || y >= ih - 2)) {
    const int c0 = in[p0];
    bool nearEdge = false;
#pragma unroll
    for (int yy = -2; yy <= 2; yy++)
#pragma unroll
      for (int xx = -2; xx <= 2; xx++) nearEdge |= in[p0 + yy * iw + xx] != c0;
    if (nearEdge) r = c0;
  }
  out[p0] = r;
}

// ---------------------------------------------------------------------------- oclrect.cl:427-464 : the vote table
// slot = ((lsid*bid) & 0x7fffffff) % nentry, no probing.  Canonical = the reference's kernel with its
// work-items in raster order: the first pixel in raster order that hits a slot claims it for its lsid, and the hit that
// performs the claim is not recorded (the atomic_cmpxchg returns 0, oclrect.cl:449-456), so the claiming pixel counts only
// if its window hits the slot a second time.  Phase 0 claims (slot word 0 = 0x7fffffff - smallest pixel index, by
// atomicMax on the cleared table), phase 1 accumulates the four maxima of the owner's hits, phase 2 replaces the claim key
// by the owner's lsid.  Keys are > any lsid (lsid < iw*ih*16/56).
#define RLS_KEY(p0) (0x7fffffff - (p0))
__device__ __forceinline__ int rls_hash(int lsid, int bid, int nentry) { return (int)((((unsigned)lsid * (unsigned)bid) & 0x7fffffffu) % (unsigned)nentry); }
// does the 7x7 window of (x, y) hit `slot` at least twice?  (only asked for the one pixel that claimed the slot, and only when
// the hit at hand has no equal right-hand neighbour in the window)
__device__ __noinline__ bool rls_hits_twice(const int *boundaryin, int lsid, int bid0, int slot, int x, int y, int iw, int ih, int nentry) {
  int c = 0;
  for (int yy = -3; yy <= 3; yy++) {
    if (y + yy < 0 || ih <= y + yy) continue;
    for (int xx = -3; xx <= 3; xx++) {
      if (x + xx < 0 || iw <= x + xx) continue;
      const int bid = boundaryin[(y + yy) * iw + x + xx];
      if (bid <= 0) continue;
      if (bid == bid0 || rls_hash(lsid, bid, nentry) == slot) { if (++c == 2) return true; }
    }
  }
  return false;
}
template <int PHASE>
__device__ __forceinline__ void rls_pixel(int *out, const int *boundaryin, const int *lsidin, int p0, int x, int y, int lsid, int iw, int ih, int nentry) {
  int lastbid = 0, twiceSlot = -1;
  bool twice = false;
  for (int yy = -3; yy <= 3; yy++) {
    if (y + yy < 0 || ih <= y + yy) continue;
    for (int xx = -3; xx <= 3; xx++) {
      if (x + xx < 0 || iw <= x + xx) continue;
      const int bid = boundaryin[(y + yy) * iw + x + xx];
      if (bid <= 0 || bid == lastbid) continue;          // consecutive repeats of a region hit the same slot with the same values
      lastbid = bid;
      const int hash = rls_hash(lsid, bid, nentry);
      int *e = out + (size_t)hash * 5;
      if (PHASE == 0) {
        atomicMax(e, RLS_KEY(p0));
      } else if (PHASE == 1) {
        const int q = RLS_KEY(e[0]);                      // the claiming pixel
        if (q != p0) { if (lsidin[q] != lsid) continue; }
        else {                                            // this pixel made the claim: it counts only if it hits the slot twice
          if (hash != twiceSlot) {
            twiceSlot = hash;
            twice = (xx < 3 && x + xx + 1 < iw && boundaryin[(y + yy) * iw + x + xx + 1] == bid) ||
                    rls_hits_twice(boundaryin, lsid, bid, hash, x, y, iw, ih, nentry);
          }
          if (!twice) continue;
        }
        // warp-aggregated: the lanes of a warp walk neighbouring string pixels, so several of them usually vote for the same
        // (segment, region) slot in the same step - they combine their boxes (redux.sync) and one lane issues the atomics
        const unsigned peers = __match_any_sync(__activemask(), hash);
        const int m1 = __reduce_max_sync(peers, iw - x), m2 = __reduce_max_sync(peers, x), m3 = __reduce_max_sync(peers, ih - y), m4 = __reduce_max_sync(peers, y);
        if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) {
          atomicMax(e + 1, m1);
          atomicMax(e + 2, m2);
          atomicMax(e + 3, m3);
          atomicMax(e + 4, m4);
        }
      } else {
        const int k = *(volatile int *)e;                 // several pixels may finalise one slot: all write the same value
        if (k > 0x3fffffff) *(volatile int *)e = lsidin[RLS_KEY(k)];
      }
    }
  }
}
template <int PHASE>
__global__ void kr_reduceLS(int *out, const int *boundaryin, const int *lsidin, int iw, int ih, int nentry, size_t fs) {
  rd_batch_z(fs, out, boundaryin, lsidin);
  XY2D;
  if (x <= 0 || y <= 0 || x >= iw - 1 || y >= ih - 1) return;
  const int lsid = lsidin[p0];
  if (lsid <= 0) return;
  rls_pixel<PHASE>(out, boundaryin, lsidin, p0, x, y, lsid, iw, ih, nentry);
}

// The same phases over the compact list of string pixels the polyline stage leaves behind (list[0] = count): only
// pixels that carry a segment id vote, and they are a few percent of the frame.
template <int PHASE>
__global__ void kr_reduceLS_list(int *out, const int *boundaryin, const int *lsidin, const int *list, int iw, int ih, int nentry, size_t fs) {
  rd_batch_y(fs, out, boundaryin, lsidin, list);
  const int count = list[0];
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
    const int p0 = list[k + 1], x = p0 % iw, y = p0 / iw;
    if (x <= 0 || y <= 0 || x >= iw - 1 || y >= ih - 1) continue;
    const int lsid = lsidin[p0];
    if (lsid <= 0) continue;
    rls_pixel<PHASE>(out, boundaryin, lsidin, p0, x, y, lsid, iw, ih, nentry);
  }
}

// ============================================================================ the oclrect_t object
// One object owns `nb` frame arenas (nb = 1 for the reference's single-frame API, oclrect.h:17-23; more for the batch
// engine).  An arena holds every device buffer of one frame with the reference's buffer plan (oclrect.c:120-135):
// buf0..5, tmp0..5, iobuf0..1 (planes of P = iw*ih*4 bytes), ioBig0..1 (4P each) and the read-back record.
#define RECT_MAGIC 0x808eae02u
struct oclrect_t {
  uint32_t magic;
  int iw, ih, ordinal, nb;
  cl_command_queue queue;
  unsigned char *dbase;                 // nb arenas, fs bytes apart
  size_t fs, P;                         // arena stride, plane pitch (bytes)
  cl_mem buf[6], tmp[6], iobuf[2], ioBig[2];   // non-owning handles on the buffers of arena 0
  cl_mem aux[7];                        // planes of the production schedule beyond the reference's plan: packed Lab (aux0), the polyline stage's
                                        // temporaries (aux1..6) - so that stage C can run beside stage B on a second stream
  cudaStream_t side;                    // the second branch of a task (stage C)
  cudaEvent_t evFork, evJoin;
  unsigned char *dblob[2];              // read-back records of arena 0, one per pipeline page: 64-byte header + rect_t list (first half), the
                                        // quadrilaterals / pose results of the device tail (second half, rd_gtail.cu).  A record may be read
                                        // - or its pose phase re-run - while the next task already runs, so the pages do not share one
  size_t blobBytes;
  int *tailTable;                       // 2 ints per region id, zero between frames (rd_gtail.cu)
  uint8_t *hostImg[2];                  // pinned: frame staging, nb x P per page (hostiobuf[page][0] of the reference)
  unsigned char *hostBlob[2];           // pinned: read-back records (first chunk of each), nb x hostBlobStride per page
  int nextPageToEnqueue, nextPageToPoll;
  cudaEvent_t events[2];
  cudaStream_t copyq;                   // second-chunk copies of long rect lists (must not queue behind the next task)
  int pending[2];                       // number of frames in flight on the page
  double pageTan[2];                    // tanAOV the pose phase of the page was enqueued with (NaN: not yet run)
  double lastTan;                       // tanAOV of the last poll: oclrect_enqueueTask has no tanAOV argument (oclrect.h:21-22), so the pose
                                        // phase runs ahead with this one and is repeated at poll time if the caller asks for another
  double wait_ms, tail_ms;              // host time spent waiting for the device / assembling the lists since the last reset
  // CUDA graphs of the single-frame task: the reference's streaming API hands over ONE frame per call, so the ~85
  // launches of a task are latency-, not throughput-bound.  One graph per page: H2D of the staged frame, the whole schedule, the
  // device tail, D2H of the record.  Captured on the second task of a page (the first one runs eagerly and sets the kernels'
  // attributes), re-captured when ws / tanAOV change.
  cudaGraphExec_t graph[2];
  double graphTan[2];
  int graphWs[2], graphLaunches[2], graphSeen[2], graphMode[2];      // graphMode: rd_get_merge_replay() at capture
};
#define FIRST_CHUNK ((size_t)16 * 1024)  // header + 90 rectangles; longer lists take a second copy

static inline int *PI(cl_mem m) { return (int *)m->dptr; }
static inline float *PF(cl_mem m) { return (float *)m->dptr; }
static inline uint32_t *PU(cl_mem m) { return (uint32_t *)m->dptr; }

__global__ void k_bgr2plab_unpack(uint32_t *plab, float *o0, float *o1, float *o2, const uint8_t *in, size_t in_fs, int iw, int ih, int ws, size_t fs);
__global__ void k_pack_plab_r(uint32_t *out, const float *i0, const float *i1, const float *i2, int n, size_t fs);
__global__ void k_edgevec_r(float2 *dst, const float *in, int iw, int ih, size_t fs);
__global__ void k_edge_plab_r(float *out, const uint32_t *in, int iw, int ih, size_t fs);
__global__ void k_thinthres_r(float *out, const float *in, const float2 *vec, int iw, int ih, size_t fs);

// executeCPUTask (oclrect.c:1049-1226) on the device.  Inputs where both schedules leave them: segment list ioBig0, region map iobuf1,
// vote table ioBig1.  Work space: the eight planes buf4 .. tmp5 (contiguous in the arena, dead by now).  The pose phase needs tanAOV;
// without one (NaN) only the quadrilaterals are prepared and the pose phase runs at poll time.
static void device_tail(oclrect_t *o, int page, double tanAOV, int nb, cudaStream_t s) {
  rd_gtail_run(o->dblob[page], o->blobBytes, (const linesegment_t *)o->ioBig[0]->dptr, PI(o->iobuf[1]), PI(o->ioBig[1]), o->tailTable,
               (unsigned char *)o->buf[4]->dptr, 8 * o->P, o->iw, o->ih, tanAOV, tanAOV == tanAOV ? 3 : 1, nb, o->fs, s);
}

// genGPUTask (oclrect.c:235-381) without the copies.  Step numbers follow SURVEY.md section 10.1; stop_step = k > 0
// returns after step k (for the intermediate-parity tests), 0 runs everything.
static void gpu_task(oclrect_t *o, const uint8_t *din, size_t din_fs, int ws, int stop_step, int nb, cudaStream_t s, int page = 0, double tanAOV = NAN) {
  const int iw = o->iw, ih = o->ih, n = iw * ih, g1 = rd_cdiv(n, 256);
  const size_t fs = o->fs;
  cl_mem *buf = o->buf, *tmp = o->tmp, *iobuf = o->iobuf, *ioBig = o->ioBig;
#define STEP(k) do { if (stop_step == (k)) return; } while (0)
  // steps 1-4 : BGR -> packed Lab (kept in buf0); recursive Gaussian r=2 on L, a, b straight from the packed plane
  // (blurred channels in tmp1..3, blurred packed Lab in buf1).  Scratch: three planes of each ioBig.
  rd_bgr2plab_run(PU(buf[0]), din, din_fs, iw, ih, ws, nb, fs, s);
  STEP(1);
  rd_iirblur3_run(PF(tmp[1]), PF(tmp[2]), PF(tmp[3]), PU(buf[1]), PU(buf[0]), PF(ioBig[1]), PF(ioBig[0]), o->P / 4, 2, iw, ih, nb, fs, s);
  STEP(3); STEP(4);
  RD_LAUNCH(k_edgevec_r, rd_gz(G2, nb), RB, 0, s, (float2 *)ioBig[0]->dptr, PF(tmp[1]), iw, ih, fs);
  STEP(5);
  RD_LAUNCH(k_edge_plab_r, rd_gz(G2, nb), RB, 0, s, PF(tmp[0]), PU(buf[1]), iw, ih, fs);
  STEP(6);
  RD_LAUNCH(k_thinthres_r, rd_gz(G2, nb), RB, 0, s, PF(buf[1]), PF(tmp[0]), (const float2 *)ioBig[0]->dptr, iw, ih, fs);
  STEP(7);
  // step 8 : edge bitmap #1
  RD_LAUNCH(kr_threshold_cast, rd_gy(g1, nb), 256, 0, s, PF(tmp[0]), PI(tmp[1]), PF(buf[1]), n, fs);
  STEP(8);
  // step 9 : junction / connect / stringify x2
  RD_LAUNCH(kr_simpleJunction, rd_gz(G2, nb), RB, 0, s, PI(buf[2]), PI(tmp[1]), iw, ih, fs);
  RD_LAUNCH(kr_simpleConnect, rd_gz(G2, nb), RB, 0, s, PI(tmp[1]), PI(buf[2]), iw, ih, fs);
  RD_LAUNCH(kr_stringify, rd_gz(G2, nb), RB, 0, s, PI(buf[2]), PI(tmp[1]), 0, iw, ih, fs);
  RD_LAUNCH(kr_stringify, rd_gz(G2, nb), RB, 0, s, PI(tmp[1]), PI(buf[2]), 1, iw, ih, fs);
  STEP(9);
  // step 10 : components of the 0/1 string image (background included, bgc = -1)
  rd_label8x(PI(buf[2]), PI(tmp[1]), tmp[0]->dptr, -1, iw, ih, nb, fs, s);
  STEP(10);
  // step 11 : strengths accumulate into buf3 (not cleared: Q1), weak components die
  RD_LAUNCH(kr_calcStrength, rd_gz(G2, nb), RB, 0, s, PI(buf[3]), PF(buf[1]), PI(buf[2]), iw, ih, fs);
  RD_LAUNCH(kr_filterStrength, rd_gz(G2, nb), RB, 0, s, PI(buf[2]), PI(buf[3]), 500, iw, ih, fs);
  STEP(11);
  // step 12 : int and i8 edge masks
  RD_LAUNCH(kr_threshold_cast_c, rd_gy(g1, nb), 256, 0, s, PI(tmp[0]), (int8_t *)tmp[1]->dptr, PI(buf[2]), n, fs);
  STEP(12);
  // step 13 : 10 x (blblur0, blblur1); walk extents live behind the i8 mask in tmp1
  rd_blblur_run(PU(buf[4]), PU(tmp[0]), PU(buf[0]), (const int8_t *)tmp[1]->dptr, (uint8_t *)tmp[1]->dptr + (size_t)n, 10, iw, ih, nb, fs, s);
  STEP(13);
  // step 14 : quantize 24^3, despeckle
  RD_LAUNCH(kr_quantize, rd_gy(g1, nb), 256, 0, s, PU(tmp[0]), PU(buf[4]), 24, 24, 24, n, fs);
  RD_LAUNCH(kr_despeckle, rd_gz(G2, nb), RB, 0, s, PU(buf[4]), PU(tmp[0]), PF(buf[1]), iw, ih, fs);
  STEP(14);
  // step 15 : strong-edge bitmap -> buf3
  RD_LAUNCH(kr_filterStrength, rd_gz(G2, nb), RB, 0, s, PI(buf[2]), PI(buf[3]), 2500, iw, ih, fs);
  RD_LAUNCH(kr_threshold_i_i, rd_gy(g1, nb), 256, 0, s, PI(buf[3]), PI(buf[2]), 0, 0, 1, n, fs);
  STEP(15);
  // step 16 : junctions of the strong edges, merge mask
  RD_LAUNCH(kr_simpleJunction, rd_gz(G2, nb), RB, 0, s, PI(tmp[0]), PI(buf[2]), iw, ih, fs);
  rd_k_clear(PI(tmp[1]), n, nb, fs, s);
  RD_LAUNCH(kr_mkMergeMask0, rd_gz(G2, nb), RB, 0, s, PI(tmp[1]), PI(tmp[0]), iw, ih, fs);
  RD_LAUNCH(kr_mkMergeMask1, rd_gz(G2, nb), RB, 0, s, PI(tmp[1]), PI(tmp[0]), iw, ih, fs);
  STEP(16);
  // step 17 : colour regions (work plane tmp4, link bytes tmp5: both dead until the polyline stage rewrites them)
  rd_labelMerge(PI(buf[5]), PI(tmp[4]), PU(buf[4]), PI(tmp[1]), PI(buf[2]), tmp[5]->dptr, PI(tmp[2]), iw, ih, nb, fs, s);     // round flags: tmp2 (dead)
  STEP(17);
  // step 18 : region sizes on top of the junction map (Q2), small regions absorbed in place in raster order (rd_despeckle2.cu;
  // input snapshot in tmp4, list planes tmp2 / tmp3 / tmp5, row counts tmp1, wide-frame row buffer ioBig0: all dead here)
  rd_calcSize_run(PI(tmp[0]), PI(buf[5]), iw, ih, nb, fs, s);
  rd_k_copy(PI(tmp[4]), PI(buf[5]), n, nb, fs, s);
  rd_despeckle2_run(PI(buf[5]), PI(tmp[4]), PI(tmp[0]), PI(tmp[2]), PI(tmp[3]), PI(tmp[5]), PI(tmp[1]), (int2 *)ioBig[0]->dptr, 16, iw, ih, nb, fs, s);
  STEP(18);
  // step 19 : boundary bands and their components -> segid map
  RD_LAUNCH(kr_markBoundary, rd_gz(G2, nb), RB, 0, s, PI(tmp[1]), PI(buf[5]), iw, ih, fs);
  rd_label8x(PI(iobuf[1]), PI(tmp[1]), tmp[0]->dptr, -1, iw, ih, nb, fs, s);
  STEP(19);
  // step 20 : polyline
  rd_polyline_run((LS_t *)ioBig[0]->dptr, n * 16, PI(buf[0]), PI(buf[3]), PI(ioBig[1]), PI(tmp[0]), PI(tmp[1]), PI(tmp[2]), PI(tmp[3]), PI(tmp[4]),
                  PI(tmp[5]), 4.0f, 20, iw, ih, nb, fs, s);
  STEP(20);
  // step 21 : vote table
  const int nentry = n * 4 / 5;
  rd_k_clear(PI(ioBig[1]), n * 4, nb, fs, s);
  RD_LAUNCH(kr_reduceLS<0>, rd_gz(G2, nb), RB, 0, s, PI(ioBig[1]), PI(iobuf[1]), PI(buf[0]), iw, ih, nentry, fs);
  RD_LAUNCH(kr_reduceLS<1>, rd_gz(G2, nb), RB, 0, s, PI(ioBig[1]), PI(iobuf[1]), PI(buf[0]), iw, ih, nentry, fs);
  RD_LAUNCH(kr_reduceLS<2>, rd_gz(G2, nb), RB, 0, s, PI(ioBig[1]), PI(iobuf[1]), PI(buf[0]), iw, ih, nentry, fs);
  STEP(21);
  // step 22 : instead of the reference's three big copies (oclrect.c:371-376) the tail runs here (scratch: buf4 .. tmp5, all dead)
  device_tail(o, page, tanAOV, nb, s);
#undef STEP
}

// The production schedule: same results as gpu_task (the step-by-step replay above, kept for the intermediate-parity
// tests), but with the fused kernels of rd_fast.cu and a buffer plan of its own.  Plane roles:
//   aux0 packed Lab   buf0 segment-id map   aux1..6 temporaries of the polyline stage   buf1 blurred packed Lab -> string labels      buf2 thinned strength
//   buf3 strength accumulator / strong-edge bitmap (carried to the next frame)      buf4 smoothed colours -> region labels
//   tmp1..3 blurred L, a, b -> weak mask + walk extents (tmp1), flat colours (tmp2), blur ping-pong / merge mask (tmp3)
//   tmp0 string bytes -> junction map + region sizes   tmp4 CCL link bytes   tmp5 strong-edge bitmap of this frame
//   iobuf1 region-boundary (segid) map   ioBig0 blur scratch -> segment list   ioBig1 blur scratch -> polyline scratch -> vote table
// stop_stage > 0 ends the schedule after that stage (tests compare the planes of the production schedule stage by stage,
// tests/parity.py FAST_STAGES); 0 runs everything.
static void gpu_task_fast(oclrect_t *o, const uint8_t *din, size_t din_fs, int ws, int nb, cudaStream_t s, int stop_stage = 0, int page = 0, double tanAOV = NAN) {
#define STAGE(k) do { if (stop_stage == (k)) return; } while (0)
  const int iw = o->iw, ih = o->ih, n = iw * ih;
  const size_t fs = o->fs;
  cl_mem *buf = o->buf, *tmp = o->tmp, *iobuf = o->iobuf, *ioBig = o->ioBig, *aux = o->aux;
  // Stage C (oclrect.c:361) needs nothing but the strong-edge bitmap of stage 6, so it runs BESIDE stages 7-12 on the object's second
  // stream (fork / join by events; inside a graph capture they become two branches).  Its clean-up kernel also copies the bitmap
  // into buf3, where the next frame's strengths accumulate.  Sequential when a test stops the schedule at a stage or
  // the per-kernel profiler is on (kernels are then timed alone).
  static const bool use_fork = !(getenv("RD_FORK") && atoi(getenv("RD_FORK")) == 0);
  const bool fork = use_fork && stop_stage == 0 && g_rd_prof_mode.load(std::memory_order_relaxed) == 0;
  auto stage_c = [&](cudaStream_t sc) {
    rd_prof_stage("C");
    rd_polyline_fast((LS_t *)ioBig[0]->dptr, n * 16, PI(buf[0]), PI(tmp[5]), PI(buf[3]), PI(ioBig[1]), PI(aux[1]), PI(aux[2]), PI(aux[3]), PI(aux[4]), PI(aux[5]),
                     PI(aux[6]), 4.0f, 20, iw, ih, nb, fs, sc);
    rd_prof_stage("B");
  };
  // Stage A (oclrect.c:245-263)
  rd_prof_stage("A");
  rd_bgr2plab_run(PU(aux[0]), din, din_fs, iw, ih, ws, nb, fs, s);
  STAGE(1);
  rd_iirblur3_run(PF(tmp[1]), PF(tmp[2]), PF(tmp[3]), PU(buf[1]), PU(aux[0]), PF(ioBig[1]), PF(ioBig[0]), o->P / 4, 2, iw, ih, nb, fs, s);
  STAGE(2);
  rd_edge_thin_run(PF(buf[2]), PF(tmp[1]), PU(buf[1]), iw, ih, nb, fs, s);
  STAGE(3);
  // Stage B (oclrect.c:265-342)
  rd_prof_stage("B");
  rd_strings1_run((uint8_t *)tmp[0]->dptr, PF(buf[2]), iw, ih, nb, fs, s);
  STAGE(4);
  rd_label8x_u8(PI(buf[1]), (const uint8_t *)tmp[0]->dptr, tmp[4]->dptr, -1, iw, ih, nb, fs, s);
  STAGE(5);
  RD_LAUNCH(kr_calcStrength, rd_gz(G2, nb), RB, 0, s, PI(buf[3]), PF(buf[2]), PI(buf[1]), iw, ih, fs);
  rd_filter_masks_run((int8_t *)tmp[1]->dptr, PI(tmp[5]), PI(buf[1]), PI(buf[3]), iw, ih, nb, fs, s);
  STAGE(6);
  if (fork) {
    RD_CUDA(cudaEventRecord(o->evFork, s));
    RD_CUDA(cudaStreamWaitEvent(o->side, o->evFork, 0));
    stage_c(o->side);
    RD_CUDA(cudaEventRecord(o->evJoin, o->side));
  }
  rd_blblur_run(PU(buf[4]), PU(tmp[3]), PU(aux[0]), (const int8_t *)tmp[1]->dptr, (uint8_t *)tmp[1]->dptr + (size_t)n, 10, iw, ih, nb, fs, s);
  STAGE(7);
  rd_quant_despeckle_run(PU(tmp[2]), PU(buf[4]), PF(buf[2]), iw, ih, nb, fs, s);
  STAGE(8);
  rd_junction_mask_run((uint8_t *)tmp[3]->dptr, PI(tmp[0]), PI(tmp[5]), iw, ih, nb, fs, s);
  STAGE(9);
  rd_labelMerge_u8(PI(buf[4]), PI(buf[5]), PU(tmp[2]), (const uint8_t *)tmp[3]->dptr, PI(tmp[5]), tmp[4]->dptr, PI(buf[1]), iw, ih, nb, fs, s);   // round flags: buf1 (dead)
  STAGE(10);
  rd_calcSize_run(PI(tmp[0]), PI(buf[4]), iw, ih, nb, fs, s);
  // despeckle2 in raster order: final labels -> buf5; list planes buf1 / buf2 / tmp2, row counts + wide-frame row buffer tmp3 (all dead here)
  rd_despeckle2_run(PI(buf[5]), PI(buf[4]), PI(tmp[0]), PI(buf[1]), PI(buf[2]), PI(tmp[2]), PI(tmp[3]), (int2 *)(PI(tmp[3]) + ((ih + 2) & ~1)), 16, iw, ih, nb, fs, s);
  rd_markBoundary_run(PI(tmp[1]), PI(buf[5]), iw, ih, nb, fs, s);
  STAGE(11);
  rd_label8x(PI(iobuf[1]), PI(tmp[1]), tmp[4]->dptr, -1, iw, ih, nb, fs, s);
  STAGE(12);
  if (fork) RD_CUDA(cudaStreamWaitEvent(s, o->evJoin, 0));
  else stage_c(s);
  STAGE(13);
  // Stage D (oclrect.c:365-367) and the device tail
  rd_prof_stage("D");
  const int nentry = n * 4 / 5;
  rd_k_clear(PI(ioBig[1]), n * 4, nb, fs, s);
  RD_LAUNCH(kr_reduceLS_list<0>, dim3(160, nb), 128, 0, s, PI(ioBig[1]), PI(iobuf[1]), PI(buf[0]), PI(aux[3]), iw, ih, nentry, fs);   // aux3: the polyline stage's pixel list
  RD_LAUNCH(kr_reduceLS_list<1>, dim3(160, nb), 128, 0, s, PI(ioBig[1]), PI(iobuf[1]), PI(buf[0]), PI(aux[3]), iw, ih, nentry, fs);
  RD_LAUNCH(kr_reduceLS_list<2>, dim3(160, nb), 128, 0, s, PI(ioBig[1]), PI(iobuf[1]), PI(buf[0]), PI(aux[3]), iw, ih, nentry, fs);
  rd_prof_stage("T");
  device_tail(o, page, tanAOV, nb, s);
}

// ---- fused / specialised Stage A kernels of the rect pipeline ----
#define RD_TABLE_QUAL static __device__ const
#include "rd_tables.inc"
__global__ void k_bgr2plab_unpack(uint32_t *plab, float *o0, float *o1, float *o2, const uint8_t *in, size_t in_fs, int iw, int ih, int ws, size_t fs) {
  rd_batch_z(fs, plab, o0, o1, o2);
  rd_batch_z(in_fs, in);
  XY2D;
  const uint8_t *p = in + (size_t)y * ws + x * 3;
  const uint32_t v = rd_srgb2plab(p[0], p[1], p[2], RD_S2L, RD_CFUNC, RD_CFUNC2);
  plab[p0] = v;
  float l, a, b;
  rd_unpacklab(v, l, a, b);
  o0[p0] = l; o1[p0] = a; o2[p0] = b;
}
__global__ void k_pack_plab_r(uint32_t *out, const float *i0, const float *i1, const float *i2, int n, size_t fs) {
  rd_batch_y(fs, out, i0, i1, i2);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = rd_packlab(i0[i], i1[i], i2[i]);
}
__global__ void k_edgevec_r(float2 *dst, const float *in, int iw, int ih, size_t fs) {
  rd_batch_z(fs, dst, in);
  XY2D;
  float vx = 0, vy = 0;
  for (int yy = -2; yy <= 2; yy++)
    for (int xx = -2; xx <= 2; xx++) {
      const float s = in[rd_mirror(x + xx, y + yy, iw, ih)];
      vx = __fadd_rn(vx, __fmul_rn(RD_V5C[(xx + 2) + (yy + 2) * 5], s));
      vy = __fadd_rn(vy, __fmul_rn(RD_V5C[(yy + 2) + (xx + 2) * 5], s));
    }
  dst[p0] = rd_edgevec_normalise(vx, vy);
}
__global__ void k_edge_plab_r(float *out, const uint32_t *in, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, in);
  XY2D;
  out[p0] = rd_edge_plab_at(in[rd_mirror(x, y - 1, iw, ih)], in[rd_mirror(x - 1, y, iw, ih)], in[rd_mirror(x, y + 1, iw, ih)],
                            in[rd_mirror(x + 1, y, iw, ih)], in[rd_mirror(x - 1, y - 1, iw, ih)], in[rd_mirror(x + 1, y + 1, iw, ih)],
                            in[rd_mirror(x + 1, y - 1, iw, ih)], in[rd_mirror(x - 1, y + 1, iw, ih)]);
}
struct RPlane {
  const float *p; int iw, ih;
  __device__ __forceinline__ float at(int x, int y) const { return p[rd_mirror(x, y, iw, ih)]; }
};
__global__ void k_thinthres_r(float *out, const float *in, const float2 *vec, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, in, vec);
  XY2D;
  RPlane pl = {in, iw, ih};
  out[p0] = rd_thinthres_at(pl, x, y, vec[p0]);
}

static void chk(oclrect_t *t) { if (!t || t->magic != RECT_MAGIC) exitf(-1, "rectdetect_b200: bad oclrect_t\n"); }

static size_t blob_need(int nrect) { return 64 + (size_t)nrect * sizeof(rect_t); }

static oclrect_t *rect_create(cl_command_queue queue, int ordinal, int iw, int ih, int nb) {
  if (rd_device_count() <= 0) exitf(-1, "rectdetect_b200: no CUDA device; there is no CPU fallback\n");
  if (iw < 8 || ih < 8) exitf(-1, "rectdetect_b200: frame %dx%d too small\n", iw, ih);
  if (!queue) exitf(-1, "rectdetect_b200: init_oclrect needs a command queue\n");
  if (nb < 1) nb = 1;
  oclrect_t *o = (oclrect_t *)calloc(1, sizeof(oclrect_t));
  o->magic = RECT_MAGIC;
  o->iw = iw; o->ih = ih; o->nb = nb;
  o->ordinal = ordinal;
  o->queue = queue;
  RD_CUDA(cudaSetDevice(o->ordinal));
  const size_t P = (((size_t)iw * ih * 4) + 255) & ~(size_t)255;
  size_t bb = (size_t)iw * ih * 2;
  if (bb < ((size_t)1 << 20)) bb = (size_t)1 << 20;
  bb = (bb + 255) & ~(size_t)255;
  o->blobBytes = bb;
  o->fs = 31 * P + 2 * bb;
  if ((iw & 3) == 0) {                                  // tensor maps (rd_tma.cuh) want the frame stride to be a multiple of the row stride
    size_t a = (size_t)iw * 4, b = 256;
    while (b) { const size_t t = a % b; a = b; b = t; }                          // a = gcd(row stride, 256)
    const size_t unit = (size_t)iw * 4 / a * 256;                                // lcm: the arenas stay 256-byte aligned
    o->fs = (o->fs + unit - 1) / unit * unit;
  }
  o->lastTan = o->pageTan[0] = o->pageTan[1] = NAN;
  o->P = P;
  RD_CUDA(cudaMalloc((void **)&o->dbase, o->fs * nb));
  // CANONICAL (Q1): memory the reference never initialises reads as zero on the first frame
  RD_CUDA(cudaMemset(o->dbase, 0, o->fs * nb));
  unsigned char *q = o->dbase;
  for (int i = 0; i < 6; i++) { o->buf[i] = rd_wrap_device_memory(q, P); q += P; }
  for (int i = 0; i < 6; i++) { o->tmp[i] = rd_wrap_device_memory(q, P); q += P; }
  for (int i = 0; i < 2; i++) { o->iobuf[i] = rd_wrap_device_memory(q, P); q += P; }
  for (int i = 0; i < 2; i++) { o->ioBig[i] = rd_wrap_device_memory(q, 4 * P); q += 4 * P; }
  for (int i = 0; i < 7; i++) { o->aux[i] = rd_wrap_device_memory(q, P); q += P; }
  o->tailTable = (int *)q; q += 2 * P;
  o->dblob[0] = q;
  o->dblob[1] = q + bb;
  RD_CUDA(cudaStreamCreateWithFlags(&o->copyq, cudaStreamNonBlocking));
  RD_CUDA(cudaStreamCreateWithFlags(&o->side, cudaStreamNonBlocking));
  RD_CUDA(cudaEventCreateWithFlags(&o->evFork, cudaEventDisableTiming));
  RD_CUDA(cudaEventCreateWithFlags(&o->evJoin, cudaEventDisableTiming));
  for (int p = 0; p < 2; p++) {
    o->hostImg[p] = (uint8_t *)allocatePinnedMemory((size_t)iw * ih * 4 * nb, NULL, NULL);
    o->hostBlob[p] = (unsigned char *)allocatePinnedMemory(FIRST_CHUNK * nb, NULL, NULL);
    RD_CUDA(cudaEventCreateWithFlags(&o->events[p], cudaEventDisableTiming | cudaEventBlockingSync));   // waiting host threads sleep: the cores run host tails
  }
  rd_quant_tables_init();
  // (the memset above ran on the legacy stream.  Not cudaDeviceSynchronize: another thread's object may be capturing its CUDA graph,
  // and a device-wide wait is not permitted while any stream captures)
  RD_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
  return o;
}

// `count` frames (<= nb), frame i at img + i*frame_stride -> device, run the schedule incl. the device tail, queue the first chunk of
// every read-back record.  src_kind 0: pageable host memory (staged through the pinned page, oclrect.c:1235/1256),
// 1: pinned host memory (copied to the device directly), 2: device memory (the schedule reads it in place).
// fresh != 0 clears the one buffer that carries state from the previous frame (buf3) so that every frame is
// processed as by a newly created object.  tanAOV = NaN: not known yet (oclrect_enqueueTask before the first poll).
static void enqueue_page(oclrect_t *o, const uint8_t *img, size_t frame_stride, int ws, int page, int src_kind, int fresh, int count, double tanAOV) {
  const int iw = o->iw, ih = o->ih;
  // ws > 0: BGR8 rows of ws bytes (the reference's format); ws < 0: an NV12 frame with a row stride of -ws bytes (Y plane, then the UV plane)
  const bool nv12 = ws < 0;
  if (!nv12 && (ws < 3 * iw || (size_t)ws * ih > (size_t)iw * ih * 4)) exitf(-1, "rectdetect_b200: row stride %d not in [3*iw, 4*iw]\n", ws);
  if (nv12 && ((iw | ih) & 1 || -ws < iw || (size_t)(-ws) * ih * 3 / 2 > (size_t)iw * ih * 4)) exitf(-1, "rectdetect_b200: NV12 needs even dimensions and a row stride in [iw, 8*iw/3] (got %d)\n", -ws);
  if (count < 1 || count > o->nb) exitf(-1, "rectdetect_b200: %d frames do not fit an object built for %d\n", count, o->nb);
  cudaStream_t s = rd_stream(o->queue);
  RD_CUDA(cudaSetDevice(o->ordinal));
  const size_t fbytes = nv12 ? (size_t)(-ws) * ih * 3 / 2 : (size_t)ws * ih, hstride = (size_t)iw * ih * 4;
  const uint8_t *din = (const uint8_t *)o->iobuf[0]->dptr;
  size_t din_fs = o->fs;
  static const bool use_graphs = !(getenv("RD_GRAPH") && atoi(getenv("RD_GRAPH")) == 0);
  const bool graphable = use_graphs && o->nb == 1 && src_kind == 0 && !fresh && g_rd_prof_mode.load(std::memory_order_relaxed) == 0;
  if (src_kind == 0)
    for (int i = 0; i < count; i++) memcpy(o->hostImg[page] + i * hstride, img + i * frame_stride, fbytes);
  if (graphable && o->graph[page] && o->graphWs[page] == ws && o->graphMode[page] == rd_get_merge_replay() && memcmp(&o->graphTan[page], &tanAOV, sizeof(double)) == 0) {
    RD_CUDA(cudaGraphLaunch(o->graph[page], s));              // its H2D node reads the staging page filled above
    g_rd_launches.fetch_add(o->graphLaunches[page], std::memory_order_relaxed);
  } else {
    const bool capture = graphable && o->graphSeen[page]++ >= 1;     // the first task of a page runs eagerly (it sets the kernels' attributes)
    int launches0 = 0;
    if (capture) {
      if (o->graph[page]) { RD_CUDA(cudaGraphExecDestroy(o->graph[page])); o->graph[page] = NULL; }
      launches0 = g_rd_launches.load(std::memory_order_relaxed);
      RD_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    }
    if (src_kind == 0) {
      RD_CUDA(cudaMemcpy2DAsync(o->iobuf[0]->dptr, o->fs, o->hostImg[page], hstride, fbytes, count, cudaMemcpyHostToDevice, s));   // oclrect.c:241
    } else if (src_kind == 1) {
      RD_CUDA(cudaMemcpy2DAsync(o->iobuf[0]->dptr, o->fs, img, count > 1 ? frame_stride : fbytes, fbytes, count, cudaMemcpyHostToDevice, s));
    } else {
      din = img;
      din_fs = frame_stride;
    }
    if (fresh) RD_CUDA(cudaMemset2DAsync(o->buf[3]->dptr, o->fs, 0, (size_t)iw * ih * 4, count, s));
    gpu_task_fast(o, din, din_fs, ws, count, s, 0, page, tanAOV);
    if (tanAOV == tanAOV) RD_CUDA(cudaMemcpy2DAsync(o->hostBlob[page], FIRST_CHUNK, o->dblob[page], o->fs, FIRST_CHUNK, count, cudaMemcpyDeviceToHost, s));
    if (capture) {
      cudaGraph_t g = NULL;
      RD_CUDA(cudaStreamEndCapture(s, &g));
      RD_CUDA(cudaGraphInstantiate(&o->graph[page], g, 0));
      RD_CUDA(cudaGraphDestroy(g));
      o->graphLaunches[page] = g_rd_launches.load(std::memory_order_relaxed) - launches0;    // (a statistic: other threads' launches may slip in)
      o->graphWs[page] = ws;
      o->graphMode[page] = rd_get_merge_replay();
      o->graphTan[page] = tanAOV;
      RD_CUDA(cudaGraphLaunch(o->graph[page], s));            // the capture recorded the task; this runs it
    }
  }
  o->pageTan[page] = tanAOV;
  RD_CUDA(cudaEventRecord(o->events[page], s));
  o->pending[page] = count;
}

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// waits for the page and turns its read-back records into rect_t lists (malloc()ed, element 0 = header, oclrect.c:1219-1225);
// out[i] receives the list of frame i.  If the pose phase has not run with this tanAOV it is run now (its inputs live in the record).
static void finish_page(oclrect_t *o, int page, double tanAOV, rect_t **out, int want_lists) {
  cudaStream_t s = rd_stream(o->queue);
  RD_CUDA(cudaSetDevice(o->ordinal));
  const double t0 = now_ms();
  const int count = o->pending[page];
  if (want_lists && memcmp(&o->pageTan[page], &tanAOV, sizeof(double)) != 0) {
    RD_CUDA(cudaStreamWaitEvent(o->copyq, o->events[page], 0));
    rd_gtail_run(o->dblob[page], o->blobBytes, NULL, NULL, NULL, NULL, NULL, 0, o->iw, o->ih, tanAOV, 2, count, o->fs, o->copyq);
    RD_CUDA(cudaMemcpy2DAsync(o->hostBlob[page], FIRST_CHUNK, o->dblob[page], o->fs, FIRST_CHUNK, count, cudaMemcpyDeviceToHost, o->copyq));
    RD_CUDA(cudaStreamSynchronize(o->copyq));
    o->pageTan[page] = tanAOV;
  } else {
    RD_CUDA(cudaEventSynchronize(o->events[page]));
  }
  o->pending[page] = 0;
  if (!want_lists) { for (int i = 0; i < count; i++) if (out) out[i] = NULL; o->wait_ms += now_ms() - t0; return; }
  const double t1 = now_ms();
  o->wait_ms += t1 - t0;
  bool more = false;
  for (int i = 0; i < count; i++) {
    const unsigned char *hb = o->hostBlob[page] + (size_t)i * FIRST_CHUNK;
    const int nrect = ((const int *)hb)[1], err = ((const int *)hb)[2];
    if (err) exitf(-1, "rectdetect_b200: the device tail ran out of %s (frame with %d line segments)\n", err == 2 ? "read-back record space" : "work space", ((const int *)hb)[0]);
    rect_t *r = (rect_t *)calloc((size_t)nrect + 1, sizeof(rect_t));
    r[0].nItems = nrect + 1;
    const size_t need = blob_need(nrect), have = need < FIRST_CHUNK ? need : FIRST_CHUNK;
    memcpy(r + 1, hb + 64, have - 64);
    if (need > FIRST_CHUNK) {                                   // long list: the rest straight from the page's device record
      RD_CUDA(cudaMemcpyAsync((unsigned char *)(r + 1) + (FIRST_CHUNK - 64), o->dblob[page] + (size_t)i * o->fs + FIRST_CHUNK, need - FIRST_CHUNK, cudaMemcpyDeviceToHost, o->copyq));
      more = true;
    }
    out[i] = r;
  }
  if (more) RD_CUDA(cudaStreamSynchronize(o->copyq));
  o->tail_ms += now_ms() - t1;
}

extern "C" {

struct oclrect_t *init_oclrect(oclimgutil_t *, oclpolyline_t *, cl_device_id device, cl_context, cl_command_queue queue, int iw, int ih) {
  return rect_create(queue, device ? device->ordinal : (queue ? queue->ordinal : 0), iw, ih, 1);
}

void dispose_oclrect(struct oclrect_t *o) {
  chk(o);
  RD_CUDA(cudaSetDevice(o->ordinal));
  RD_CUDA(cudaStreamSynchronize(rd_stream(o->queue)));
  for (int i = 0; i < 6; i++) { clReleaseMemObject(o->buf[i]); clReleaseMemObject(o->tmp[i]); }
  for (int i = 0; i < 2; i++) { clReleaseMemObject(o->iobuf[i]); clReleaseMemObject(o->ioBig[i]); }
  RD_CUDA(cudaFree(o->dbase));
  for (int p = 0; p < 2; p++) { freePinnedMemory(o->hostImg[p], NULL, NULL); freePinnedMemory(o->hostBlob[p], NULL, NULL); RD_CUDA(cudaEventDestroy(o->events[p])); }
  RD_CUDA(cudaStreamDestroy(o->copyq));
  RD_CUDA(cudaStreamSynchronize(o->side));
  RD_CUDA(cudaStreamDestroy(o->side));
  RD_CUDA(cudaEventDestroy(o->evFork));
  RD_CUDA(cudaEventDestroy(o->evJoin));
  for (int i = 0; i < 7; i++) clReleaseMemObject(o->aux[i]);
  for (int p = 0; p < 2; p++) if (o->graph[p]) RD_CUDA(cudaGraphExecDestroy(o->graph[p]));
  o->magic = 0;
  free(o);
}

rect_t *oclrect_executeOnce(struct oclrect_t *o, uint8_t *imgData, int ws, const double tanAOV) {   // oclrect.c:1230
  chk(o);
  if (o->pending[0]) exitf(-1, "rectdetect_b200: oclrect_executeOnce while a task is pending on page 0\n");
  enqueue_page(o, imgData, 0, ws, 0, 0, 0, 1, tanAOV);
  o->lastTan = tanAOV;
  rect_t *r = NULL;
  finish_page(o, 0, tanAOV, &r, 1);
  return r;
}

void oclrect_enqueueTask(struct oclrect_t *o, uint8_t *imgData, int ws) {                            // oclrect.c:1248
  chk(o);
  const int page = 1 & o->nextPageToEnqueue;
  o->nextPageToEnqueue++;
  if (o->pending[page]) exitf(-1, "rectdetect_b200: oclrect_enqueueTask with two tasks already in flight\n");   // assert(events[page]==NULL)
  enqueue_page(o, imgData, 0, ws, page, 0, 0, 1, o->lastTan);
}

rect_t *oclrect_pollTask(struct oclrect_t *o, const double tanAOV) {                                // oclrect.c:1263
  chk(o);
  const int page = 1 & o->nextPageToPoll;
  o->nextPageToPoll++;
  if (!o->pending[page]) exitf(-1, "rectdetect_b200: oclrect_pollTask without a pending task\n");
  rect_t *r = NULL;
  o->lastTan = tanAOV;
  finish_page(o, page, tanAOV, &r, 1);
  return r;
}

cl_mem rd_oclrect_buffer(struct oclrect_t *o, const char *name) {
  chk(o);
  if (!strncmp(name, "buf", 3) && name[3] >= '0' && name[3] < '6') return o->buf[name[3] - '0'];
  if (!strncmp(name, "tmp", 3) && name[3] >= '0' && name[3] < '6') return o->tmp[name[3] - '0'];
  if (!strncmp(name, "iobuf", 5) && name[5] >= '0' && name[5] < '2') return o->iobuf[name[5] - '0'];
  if (!strncmp(name, "ioBig", 5) && name[5] >= '0' && name[5] < '2') return o->ioBig[name[5] - '0'];
  if (!strncmp(name, "aux", 3) && name[3] >= '0' && name[3] < '7') return o->aux[name[3] - '0'];
  return NULL;
}

void rd_oclrect_run_device(struct oclrect_t *o, const uint8_t *imgData, int ws, int stop_step) {
  chk(o);
  cudaStream_t s = rd_stream(o->queue);
  RD_CUDA(cudaMemcpyAsync(o->iobuf[0]->dptr, imgData, (size_t)ws * o->ih, cudaMemcpyHostToDevice, s));
  if (stop_step > 0) gpu_task(o, (const uint8_t *)o->iobuf[0]->dptr, o->fs, ws, stop_step, 1, s);
  else gpu_task_fast(o, (const uint8_t *)o->iobuf[0]->dptr, o->fs, ws, 1, s, -stop_step);      // stop_step < 0: stage -stop_step of the production schedule
  RD_CUDA(cudaStreamSynchronize(s));
}

// ---- Stage B / D operators, one per __kernel of oclrect.cl ----
#define QS cudaStream_t s = rd_stream(q); const int nb = 1; const size_t fs = 0
void rd_rect_simpleJunction(cl_mem out, cl_mem in, int iw, int ih, cl_command_queue q) { QS; RD_LAUNCH(kr_simpleJunction, rd_gz(G2, nb), RB, 0, s, PI(out), PI(in), iw, ih, fs); }
void rd_rect_simpleConnect(cl_mem out, cl_mem in, int iw, int ih, cl_command_queue q) { QS; RD_LAUNCH(kr_simpleConnect, rd_gz(G2, nb), RB, 0, s, PI(out), PI(in), iw, ih, fs); }
void rd_rect_stringify(cl_mem out, cl_mem in, int mod2, int iw, int ih, cl_command_queue q) { QS; RD_LAUNCH(kr_stringify, rd_gz(G2, nb), RB, 0, s, PI(out), PI(in), mod2, iw, ih, fs); }
void rd_rect_blblur0(cl_mem out, cl_mem e, cl_mem in, int iw, int ih, cl_command_queue q) { QS; RD_LAUNCH(kr_blblur<0>, rd_gz(G2, nb), RB, 0, s, PU(out), (const int8_t *)e->dptr, PU(in), iw, ih, fs); }
void rd_rect_blblur1(cl_mem out, cl_mem e, cl_mem in, int iw, int ih, cl_command_queue q) { QS; RD_LAUNCH(kr_blblur<1>, rd_gz(G2, nb), RB, 0, s, PU(out), (const int8_t *)e->dptr, PU(in), iw, ih, fs); }
void rd_rect_quantize(cl_mem out, cl_mem in, int n0, int n1, int n2, int iw, int ih, cl_command_queue q) { QS; RD_LAUNCH(kr_quantize, rd_gy(rd_cdiv(iw * ih, 256), nb), 256, 0, s, PU(out), PU(in), n0, n1, n2, iw * ih, fs); }
void rd_rect_despeckle(cl_mem out, cl_mem in, cl_mem edge, int iw, int ih, cl_command_queue q) { QS; RD_LAUNCH(kr_despeckle, rd_gz(G2, nb), RB, 0, s, PU(out), PU(in), PF(edge), iw, ih, fs); }
void rd_rect_mkMergeMask0(cl_mem out, cl_mem j, int iw, int ih, cl_command_queue q) { QS; RD_LAUNCH(kr_mkMergeMask0, rd_gz(G2, nb), RB, 0, s, PI(out), PI(j), iw, ih, fs); }
void rd_rect_mkMergeMask1(cl_mem io, cl_mem j, int iw, int ih, cl_command_queue q) { QS; RD_LAUNCH(kr_mkMergeMask1, rd_gz(G2, nb), RB, 0, s, PI(io), PI(j), iw, ih, fs); }
void rd_rect_labelMerge(cl_mem label, cl_mem pix, cl_mem mask, cl_mem edge, int iw, int ih, cl_command_queue q) {
  QS;
  int *work = NULL; void *links = NULL;
  RD_CUDA(cudaMallocAsync((void **)&work, (size_t)iw * ih * 4 + 64, s));
  RD_CUDA(cudaMallocAsync(&links, (size_t)iw * ih * 4, s));
  rd_labelMerge(PI(label), work, PU(pix), PI(mask), PI(edge), links, work + (size_t)iw * ih, iw, ih, 1, 0, s);
  RD_CUDA(cudaFreeAsync(work, s));
  RD_CUDA(cudaFreeAsync(links, s));
}
void rd_rect_calcSize(cl_mem out, cl_mem label, int iw, int ih, cl_command_queue q) { QS; RD_LAUNCH(kr_calcSize, rd_gy(rd_cdiv(iw * ih, 256), nb), 256, 0, s, PI(out), PI(label), iw * ih, fs); }
void rd_rect_despeckle2(cl_mem io, cl_mem size, cl_mem scratch, int thre, int iw, int ih, cl_command_queue q) {
  QS;
  const size_t n = (size_t)iw * ih;
  int *w = NULL;                                   // three list planes, the row counts, the wide-frame row buffer
  RD_CUDA(cudaMallocAsync((void **)&w, (3 * n + ih + 4 * (size_t)iw + 4) * sizeof(int), s));
  rd_k_copy(PI(scratch), PI(io), iw * ih, 1, 0, s);
  int *rowbuf = w + 3 * n + ih;
  rowbuf += ((uintptr_t)rowbuf & 4) ? 1 : 0;
  rd_despeckle2_run(PI(io), PI(scratch), PI(size), w, w + n, w + 2 * n, w + 3 * n, (int2 *)rowbuf, thre, iw, ih, nb, fs, s);
  RD_CUDA(cudaFreeAsync(w, s));
}
// executeCPUTask (oclrect.c:1049) on caller-owned device buffers: the device tail as an operator.  Returns a malloc()ed list.
rect_t *rd_rect_tail_device(cl_mem lsList, cl_mem segid, cl_mem votes, int iw, int ih, double tanAOV, cl_command_queue q) {
  cudaStream_t s = rd_stream(q);
  const size_t P = (((size_t)iw * ih * 4) + 255) & ~(size_t)255;
  size_t bb = (size_t)iw * ih * 2;
  if (bb < ((size_t)1 << 20)) bb = (size_t)1 << 20;
  unsigned char *w = NULL;                         // table (2P, zero), work space (8P), record
  RD_CUDA(cudaMallocAsync((void **)&w, 10 * P + bb, s));
  RD_CUDA(cudaMemsetAsync(w, 0, 2 * P, s));
  unsigned char *blob = w + 10 * P;
  rd_gtail_run(blob, bb, (const linesegment_t *)lsList->dptr, PI(segid), PI(votes), (int *)w, w + 2 * P, 8 * P, iw, ih, tanAOV, 3, 1, 0, s);
  int hdr[16];
  RD_CUDA(cudaMemcpyAsync(hdr, blob, sizeof(hdr), cudaMemcpyDeviceToHost, s));
  RD_CUDA(cudaStreamSynchronize(s));
  if (hdr[2]) exitf(-1, "rectdetect_b200: the device tail ran out of %s (frame with %d line segments)\n", hdr[2] == 2 ? "read-back record space" : "work space", hdr[0]);
  rect_t *r = (rect_t *)calloc((size_t)hdr[1] + 1, sizeof(rect_t));
  r[0].nItems = hdr[1] + 1;
  if (hdr[1] > 0) RD_CUDA(cudaMemcpyAsync(r + 1, blob + 64, (size_t)hdr[1] * sizeof(rect_t), cudaMemcpyDeviceToHost, s));
  RD_CUDA(cudaFreeAsync(w, s));
  RD_CUDA(cudaStreamSynchronize(s));
  return r;
}
void rd_rect_markBoundary(cl_mem out, cl_mem in, int iw, int ih, cl_command_queue q) { QS; RD_LAUNCH(kr_markBoundary, rd_gz(G2, nb), RB, 0, s, PI(out), PI(in), iw, ih, fs); }
void rd_rect_reduceLS(cl_mem out, cl_mem boundary, cl_mem lsid, int iw, int ih, int nentry, cl_command_queue q) {
  QS;
  RD_LAUNCH(kr_reduceLS<0>, rd_gz(G2, nb), RB, 0, s, PI(out), PI(boundary), PI(lsid), iw, ih, nentry, fs);
  RD_LAUNCH(kr_reduceLS<1>, rd_gz(G2, nb), RB, 0, s, PI(out), PI(boundary), PI(lsid), iw, ih, nentry, fs);
  RD_LAUNCH(kr_reduceLS<2>, rd_gz(G2, nb), RB, 0, s, PI(out), PI(boundary), PI(lsid), iw, ih, nentry, fs);
}



// ============================================================================ frame-batch engine
// nctx pipeline objects, each with its own stream and room for `fpl` frames per launch, each driven by one host
// thread.  Every kernel launch processes a chunk of up to fpl independent frames (blockIdx.z / .y = frame), which
// amortises launch latency and fills the 148 SMs even where one frame offers little parallelism (the recursive-
// Gaussian scans have 2 chains per row).  A worker keeps two chunks in flight on its object (enqueue N+1 before
// polling N, vidrect.cpp:159-172), so the device stages overlap the host tails, and the nctx streams overlap each other.
}  // extern "C" (reopened below)

struct rd_batch {
  int device, iw, ih, nctx, fpl;
  std::vector<oclrect_t *> ctx;
  std::vector<cl_command_queue> queues;
  double stage_ms[5];
};

static int host_ptr_kind(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return 2;
  if (a.type == cudaMemoryTypeHost) return 1;
  return 0;
}

static void batch_worker(rd_batch *b, int c, const uint8_t *frames, size_t frame_stride, int ws, int nframes, double tanAOV, rect_t **out, int kind, int run_tail) {
  oclrect_t *o = b->ctx[c];
  RD_CUDA(cudaSetDevice(b->device));
  const int nchunks = (nframes + b->fpl - 1) / b->fpl;
  int prev = -1, prevPage = 0, page = 0;
  std::vector<rect_t *> tmp(b->fpl);
  auto finish = [&](int chunk, int pg) {
    const int f0 = chunk * b->fpl;
    finish_page(o, pg, tanAOV, tmp.data(), run_tail);
    const int cnt = (nframes - f0) < b->fpl ? (nframes - f0) : b->fpl;
    for (int i = 0; i < cnt; i++) { if (out) out[f0 + i] = tmp[i]; else free(tmp[i]); }
  };
  for (int ch = c; ch < nchunks; ch += b->nctx) {
    const int f0 = ch * b->fpl;
    const int cnt = (nframes - f0) < b->fpl ? (nframes - f0) : b->fpl;
    enqueue_page(o, frames + (size_t)f0 * frame_stride, frame_stride, ws, page, kind, 1, cnt, tanAOV);
    if (prev >= 0) finish(prev, prevPage);
    prev = ch; prevPage = page; page ^= 1;
  }
  if (prev >= 0) finish(prev, prevPage);
}

extern "C" {

rd_batch *rd_batch_create(int device, int iw, int ih, int nctx, int frames_per_launch) {
  if (rd_device_count() <= device || device < 0) exitf(-1, "rectdetect_b200: rd_batch_create: no CUDA device %d; there is no CPU fallback\n", device);
  if (nctx < 1) nctx = 1;
  if (frames_per_launch < 1) frames_per_launch = 1;
  rd_batch *b = new rd_batch();
  b->device = device; b->iw = iw; b->ih = ih; b->nctx = nctx; b->fpl = frames_per_launch;
  cl_device_id dev = simpleGetDevice(device);
  for (int c = 0; c < nctx; c++) {
    cl_command_queue q = clCreateCommandQueue(NULL, dev, 0, NULL);
    b->queues.push_back(q);
    b->ctx.push_back(rect_create(q, device, iw, ih, frames_per_launch));
  }
  for (int i = 0; i < 5; i++) b->stage_ms[i] = 0;
  return b;
}

void rd_batch_destroy(rd_batch *b) {
  if (!b) return;
  for (int c = 0; c < b->nctx; c++) { dispose_oclrect(b->ctx[c]); clReleaseCommandQueue(b->queues[c]); }
  delete b;
}

static void batch_run(rd_batch *b, const uint8_t *frames, size_t frame_stride, int ws, int nframes, double tanAOV, rect_t **out, int kind, int run_tail) {
  for (oclrect_t *o : b->ctx) { o->wait_ms = 0; o->tail_ms = 0; }
  const int nchunks = (nframes + b->fpl - 1) / b->fpl;
  const int nw = b->nctx < nchunks ? b->nctx : nchunks;
  std::vector<std::thread> th;
  for (int c = 1; c < nw; c++) th.emplace_back(batch_worker, b, c, frames, frame_stride, ws, nframes, tanAOV, out, kind, run_tail);
  if (nw > 0) batch_worker(b, 0, frames, frame_stride, ws, nframes, tanAOV, out, kind, run_tail);
  for (auto &t : th) t.join();
}

void rd_batch_run(rd_batch *b, const uint8_t *frames, size_t frame_stride, int ws, int nframes, double tanAOV, rect_t **out) {
  int kind = host_ptr_kind(frames);
  if (kind == 2) exitf(-1, "rectdetect_b200: rd_batch_run expects host frames (use rd_batch_run_device)\n");
  batch_run(b, frames, frame_stride, ws, nframes, tanAOV, out, kind, 1);
}

// NV12 frames (Y plane with row stride ystride, UV plane behind it), in host (pageable or pinned) or device memory
void rd_batch_run_nv12(rd_batch *b, const void *frames, size_t frame_stride, int ystride, int nframes, double tanAOV, rect_t **out) {
  if (ystride <= 0) exitf(-1, "rectdetect_b200: rd_batch_run_nv12: bad row stride %d\n", ystride);
  batch_run(b, (const uint8_t *)frames, frame_stride, -ystride, nframes, tanAOV, out, host_ptr_kind(frames), 1);
}
rect_t *rd_oclrect_executeOnceNV12(struct oclrect_t *o, const uint8_t *nv12, int ystride, double tanAOV) {
  chk(o);
  if (o->pending[0]) exitf(-1, "rectdetect_b200: rd_oclrect_executeOnceNV12 while a task is pending on page 0\n");
  if (ystride <= 0) exitf(-1, "rectdetect_b200: rd_oclrect_executeOnceNV12: bad row stride %d\n", ystride);
  enqueue_page(o, nv12, 0, -ystride, 0, 0, 0, 1, tanAOV);
  o->lastTan = tanAOV;
  rect_t *r = NULL;
  finish_page(o, 0, tanAOV, &r, 1);
  return r;
}

void rd_batch_run_device(rd_batch *b, const void *dframes, size_t frame_stride, int ws, int nframes, double tanAOV, rect_t **out) {
  batch_run(b, (const uint8_t *)dframes, frame_stride, ws, nframes, tanAOV, out, 2, out != NULL);
}

// n rect_t lists (as the run functions return them) -> one malloc()ed array of all their entries, headers dropped, in list order;
// counts[i] = entries of list i.  The lists are freed.  (For callers that handle thousands of lists per second: bench.py, the gather.)
rect_t *rd_rect_lists_flatten(rect_t **lists, int n, int32_t *counts) {
  size_t total = 0;
  for (int i = 0; i < n; i++) { counts[i] = lists[i] ? lists[i][0].nItems - 1 : 0; total += (size_t)counts[i]; }
  rect_t *flat = (rect_t *)malloc((total ? total : 1) * sizeof(rect_t));
  size_t o = 0;
  for (int i = 0; i < n; i++) {
    if (counts[i] > 0) memcpy(flat + o, lists[i] + 1, (size_t)counts[i] * sizeof(rect_t));
    o += (size_t)counts[i];
    free(lists[i]);
    lists[i] = NULL;
  }
  return flat;
}

void rd_batch_stage_ms(rd_batch *b, double out_ms[5]) {
  for (int i = 0; i < 5; i++) out_ms[i] = 0;
  for (oclrect_t *o : b->ctx) { out_ms[0] += o->wait_ms; out_ms[4] += o->tail_ms; }
}

}  // extern "C"
