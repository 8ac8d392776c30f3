// rd_fast.cu - the tuned kernels of the rect pipeline's device schedule (rd_rect.cu : gpu_task).
//
// The operator-level kernels in rd_imgutil.cu / rd_rect.cu follow the reference one launch per __kernel.  The
// schedule replaces the expensive groups by the kernels below, which compute bit-identical results:
//   rd_blblur_run     : the 10 x (blblur0, blblur1) edge-stopped box blurs (oclrect.cl:155-205, oclrect.c:286-296).
//                       The stop rules only look at the edge mask, which is the same for all 20 passes, so the walk
//                       extents of every pixel are computed once (1 B/px per direction) and each iteration becomes
//                       one shared-memory-tiled kernel doing the x pass and the y pass back to back.
//   rd_calcSize_run   : region histogram (oclrect.cl:336) with warp- and CTA-level aggregation in front of the
//                       global atomics (one region can own half a frame).
//   rd_iirblur3_run   : the three recursive-Gaussian blurs of oclrect.c:248-250 (oclimgutil.cl:542-637) straight
//                       from the packed Lab plane: one thread runs the three channel chains of a row (column)
//                       and direction, the pass1 / pass3 combinations are folded into the consumers.
#include "rd_common.cuh"
#include "rd_tma.cuh"
#include "rd_stageA.cuh"
#include "rd_bits.cuh"
#include <mutex>
#define RD_TABLE_QUAL static __device__ const
#include "rd_tables.inc"

// =============================================================================================== blblur
#define BLB 4
// walk extents of pixel (pos along the walk, p0) : nl = pixels taken walking back (0..5), nr = walking forward (0..5)
template <int DIR, class E>
__device__ __forceinline__ unsigned blb_extent(const E &e, int x, int y, int iw, int ih) {
  const int pos = DIR == 0 ? x : y, len = DIR == 0 ? iw : ih;
  const bool hasSide = DIR == 0 ? (y < ih - 1) : (x < iw - 1);
  const int oe = e.at(x, y) != 0;
  int nl = 0, nr = 0;
  for (int d = 0; d >= -BLB; d--) {
    if (pos + d < 0) break;
    const int qx = DIR == 0 ? x + d : x, qy = DIR == 0 ? y : y + d;
    const int bx = DIR == 0 ? qx - 1 : qx, by = DIR == 0 ? qy : qy - 1;          // previous pixel of the walk
    const int sx = DIR == 0 ? qx : qx + 1, sy = DIR == 0 ? qy + 1 : qy;          // perpendicular neighbour
    if (pos + d > 0 && e.at(qx, qy) != 0 && e.at(bx, by) == 0) break;
    if (pos + d > 0 && hasSide && e.at(qx, qy) == 0 && e.at(bx, by) != 0 && e.at(sx, sy) != 0) break;
    nl++;
  }
  for (int d = 0; d <= BLB; d++) {
    if (pos + d > len - 1) break;
    const int qx = DIR == 0 ? x + d : x, qy = DIR == 0 ? y : y + d;
    const int fx = DIR == 0 ? qx + 1 : qx, fy = DIR == 0 ? qy : qy + 1;          // next pixel of the walk
    if (pos + d < len - 1 && e.at(qx, qy) == 0 && e.at(fx, fy) != 0) break;
    if (oe && e.at(qx, qy) == 0) break;
    nr++;
  }
  return (unsigned)nl | ((unsigned)nr << 4);
}

#define EX_TX 64
#define EX_TY 16
#define EX_H 5
struct SmemEdge {
  const int8_t *t; int x0, y0;           // tile origin (image coords of t[0]) ; row pitch EX_TX + 2*EX_H
  __device__ __forceinline__ int at(int x, int y) const { return t[(y - y0) * (EX_TX + 2 * EX_H) + (x - x0)]; }
};
__global__ void __launch_bounds__(256) kf_blb_extents(uint8_t *extH, uint8_t *extV, const int8_t *edge, int iw, int ih, size_t fs) {
  rd_batch_z(fs, extH, extV, edge);
  __shared__ int8_t tile[(EX_TY + 2 * EX_H) * (EX_TX + 2 * EX_H)];
  const int x0 = blockIdx.x * EX_TX - EX_H, y0 = blockIdx.y * EX_TY - EX_H;
  for (int i = threadIdx.x; i < (EX_TY + 2 * EX_H) * (EX_TX + 2 * EX_H); i += 256) {
    const int tx = i % (EX_TX + 2 * EX_H), ty = i / (EX_TX + 2 * EX_H);
    const int gx = x0 + tx, gy = y0 + ty;
    tile[i] = (gx >= 0 && gx < iw && gy >= 0 && gy < ih) ? edge[(size_t)gy * iw + gx] : (int8_t)0;   // never read when outside
  }
  __syncthreads();
  SmemEdge e = {tile, x0, y0};
#pragma unroll
  for (int k = 0; k < (EX_TX * EX_TY) / 256; k++) {
    const int i = threadIdx.x + k * 256;
    const int x = blockIdx.x * EX_TX + (i % EX_TX), y = blockIdx.y * EX_TY + (i / EX_TX);
    if (x < iw && y < ih) {
      extH[(size_t)y * iw + x] = (uint8_t)blb_extent<0>(e, x, y, iw, ih);
      extV[(size_t)y * iw + x] = (uint8_t)blb_extent<1>(e, x, y, iw, ih);
    }
  }
}

// floor(c / w) for 0 <= c < 2^16, 1 <= w <= 10 : one multiply-high by ceil(2^32 / w)
__constant__ const unsigned BLB_RCP[11] = {0u, 0u, 0x80000000u, 0x55555556u, 0x40000000u, 0x33333334u, 0x2aaaaaabu, 0x24924925u, 0x20000000u, 0x1c71c71du, 0x1999999au};
__device__ __forceinline__ unsigned blb_div(unsigned c, unsigned w) { return w == 1 ? c : __umulhi(c, BLB_RCP[w]); }
__device__ __forceinline__ unsigned blb_div_r(unsigned c, unsigned rcp) { return rcp == 0 ? c : __umulhi(c, rcp); }
__device__ __forceinline__ uint32_t blb_mean(unsigned c0, unsigned c1, unsigned c2, unsigned w) {
  // packlabbl(csum / wsum): the quotients cannot leave their fields (means of in-range values), so no clamp is needed
  return (blb_div(c2, w) << 22) | (blb_div(c1, w) << 12) | blb_div(c0, w);
}

// Shared-memory tiles hold the three channels spread out in one 64-bit word (L at bit 0, a at bit 20, b at bit 40), so
// a walk is a chain of plain 64-bit adds (at most 10 terms of at most 12 bits each: no field can overflow into the next).
typedef unsigned long long blb_w;
__device__ __forceinline__ blb_w blb_spread(uint32_t v) { return (blb_w)(v & 4095u) | ((blb_w)((v >> 12) & 1023u) << 20) | ((blb_w)(v >> 22) << 40); }
__device__ __forceinline__ uint32_t blb_pack(blb_w w) { return (uint32_t)(w & 4095u) | ((uint32_t)((w >> 20) & 1023u) << 12) | ((uint32_t)(w >> 40) << 22); }
__device__ __forceinline__ uint32_t blb_finish(blb_w acc, blb_w centre, unsigned w) {
  if (w == 0) return blb_pack(centre);
  return blb_mean((unsigned)(acc & 0xfffffu), (unsigned)((acc >> 20) & 0xfffffu), (unsigned)(acc >> 40), w);
}

// ---- the iteration kernel -------------------------------------------------------------------------------------------
// (Two CTA-tiled variants were measured first - a 64x32 tile staged by plain loads, and a persistent one staged by
// cp.async.bulk into a double buffer - and both sat at ~15 us per frame-iteration on CTA barriers and apron recomputation;
// see DESIGN.md.)
// One WARP owns a 32-pixel-wide strip of up to SB_CH rows and streams down it, so nothing ever waits on a CTA barrier:
//  - per row it loads 40 pixels (apron 4 each side) and the row's x-extents, forms the running sums of the row with a
//    shuffle scan, and every lane takes its x-pass mean from four entries of that row of sums;
//  - each lane keeps the running sum of ITS column of x-pass results in a 16-row ring in shared memory, and as soon as
//    row y+4 is in, the y-pass mean of row y is four ring entries away.  No row is loaded or x-filtered twice within a strip
//    (the tiled kernels above redo the 8 apron rows of every 32-row tile).
// Field widths: a strip adds at most SB_CH_MAX + 8 values of at most 12 bits per field, inside the 20-bit fields.
#define SB_CH_MIN 32                 // strip heights are chosen per launch (see rd_blblur_run) within [SB_CH_MIN, SB_CH_MAX]
#define SB_CH_MAX 240                // (240 + 8) rows x 4095 still fits the 20-bit fields of the running sums
#define SB_RING 16
#define SB_WARPS 8
__device__ __forceinline__ blb_w blb_scan_up(blb_w v, int lane, int steps) {
#pragma unroll
  for (int k = 0; k < 5; k++) {
    if (k < steps) { const blb_w t = __shfl_up_sync(0xffffffffu, v, 1 << k); if (lane >= (1 << k)) v += t; }
  }
  return v;
}
__global__ void __launch_bounds__(SB_WARPS * 32) kf_blb_stream(uint32_t *out, const uint32_t *in, const uint8_t *extH, const uint8_t *extV, int iw, int ih, int nb,
                                                               int strips, int chunks, int ch, size_t fs) {
  __shared__ blb_w rowP[SB_WARPS][48];
  __shared__ blb_w ring[SB_WARPS][SB_RING][32];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int wid = blockIdx.x * SB_WARPS + wp;
  if (wid >= strips * chunks * nb) return;
  const int sx = wid % strips, cy = (wid / strips) % chunks, z = wid / (strips * chunks);
  rd_batch_off((size_t)z * fs, out, in, extH, extV);
  const int x = sx * 32 + lane, xa = x - BLB, xb = x + 32 - BLB;
  const bool oka = xa >= 0 && xa < iw, okb = lane < 2 * BLB && xb < iw, okx = x < iw;
  const int y0 = cy * ch, y1 = min(y0 + ch, ih);
  const int first = max(y0 - BLB, 0), last = min(y1 + BLB, ih);
  blb_w *P = rowP[wp];
  blb_w (*R)[32] = ring[wp];
  blb_w Q = 0;
  R[first & (SB_RING - 1)][lane] = 0;
  // two rows of look-ahead in registers
  uint32_t a0 = 0, b0 = 0, a1 = 0, b1 = 0;
  unsigned e0 = 0, e1 = 0;
  {
    const size_t r0 = (size_t)first * iw, r1 = (size_t)(first + 1) * iw;
    if (oka) a0 = in[r0 + xa];
    if (okb) b0 = in[r0 + xb];
    if (okx) e0 = extH[r0 + x];
    if (first + 1 < last) { if (oka) a1 = in[r1 + xa]; if (okb) b1 = in[r1 + xb]; if (okx) e1 = extH[r1 + x]; }
  }
  for (int y = first; y < last; y++) {
    const uint32_t a = a0, b = b0;
    const unsigned e = e0;
    a0 = a1; b0 = b1; e0 = e1;
    a1 = 0; b1 = 0; e1 = 0;
    if (y + 2 < last) {
      const size_t r2 = (size_t)(y + 2) * iw;
      if (oka) a1 = in[r2 + xa];
      if (okb) b1 = in[r2 + xb];
      if (okx) e1 = extH[r2 + x];
    }
    const int yv = y - BLB;
    unsigned ev = 0;
    const bool doV = yv >= y0 && okx;
    if (doV) ev = extV[(size_t)yv * iw + x];
    // x pass: running sums of the 40-pixel row
    const blb_w wa = blb_spread(a), wb = blb_spread(b);
    const blb_w ia = blb_scan_up(wa, lane, 5);
    const blb_w ta = __shfl_sync(0xffffffffu, ia, 31);
    const blb_w ib = blb_scan_up(wb, lane, 3);
    P[lane] = ia - wa;
    if (lane < 2 * BLB) P[32 + lane] = ta + ib - wb;
    if (lane == 2 * BLB - 1) P[32 + 2 * BLB] = ta + ib;
    __syncwarp();
    const int c = lane + BLB, nl = e & 15, nr = e >> 4;
    const blb_w pc = P[c], pc1 = P[c + 1];
    const blb_w hs = (pc1 - P[c + 1 - nl]) + (P[c + nr] - pc);
    const blb_w h = (nl + nr) == 0 ? (pc1 - pc) : blb_spread(blb_finish(hs, 0, nl + nr));
    __syncwarp();
    // y pass: column running sums in the ring; row yv = y - 4 is complete once Q[y + 1] is known
    Q += h;
    R[(y + 1) & (SB_RING - 1)][lane] = Q;
    if (doV) {
      const int nu = ev & 15, nd = ev >> 4;
      const blb_w qa = R[(yv + 1) & (SB_RING - 1)][lane], qd = R[yv & (SB_RING - 1)][lane];
      const blb_w vs = (qa - R[(yv + 1 - nu) & (SB_RING - 1)][lane]) + (R[(yv + nd) & (SB_RING - 1)][lane] - qd);
      out[(size_t)yv * iw + x] = blb_finish(vs, qa - qd, nu + nd);
    }
  }
  // rows whose downward walk is cut by the bottom of the image rather than by the strip
  if (okx) {
    for (int yv = max(y0, last - BLB); yv < y1; yv++) {
      const unsigned ev = extV[(size_t)yv * iw + x];
      const int nu = ev & 15, nd = ev >> 4;
      const blb_w qa = R[(yv + 1) & (SB_RING - 1)][lane], qd = R[yv & (SB_RING - 1)][lane];
      const blb_w vs = (qa - R[(yv + 1 - nu) & (SB_RING - 1)][lane]) + (R[(yv + nd) & (SB_RING - 1)][lane] - qd);
      out[(size_t)yv * iw + x] = blb_finish(vs, qa - qd, nu + nd);
    }
  }
}

// ---- walk extents, bit-parallel ------------------------------------------------------------------------------------
// The stop rules of blb_extent() only combine the edge bit of a position with those of the previous / next position of
// the walk and of the perpendicular neighbour, so they can be evaluated for 32 positions at once on bit rows:
//   stopBack(q) = E(q) & ~E(q-1)  |  hasSide & ~E(q) & E(q-1) & Side(q)        (never at q == 0; q < 0 always stops)
//   stopFwd(q)  = ~E(q) & E(q+1)  |  oe & ~E(q)                                 (q beyond the last position always stops)
// and an extent is the number of clear bits in front of the first stop (at most 5).  One warp streams down a strip
// of 32 columns: the rows arrive as ballot words (x walk), every lane keeps the history of its own column and of the
// column to its right in two registers (y walk: row y - b is bit b).
#define EXS_WARPS 8
__device__ __forceinline__ unsigned exs_stop_back(unsigned e, unsigned eprev, unsigned side) { return (e & ~eprev) | (~e & eprev & side); }
__global__ void __launch_bounds__(EXS_WARPS * 32) kf_blb_extents_s(uint8_t *extH, uint8_t *extV, const int8_t *edge, int iw, int ih, int nb, int strips, int chunks,
                                                                   int ch, size_t fs) {
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * EXS_WARPS + (threadIdx.x >> 5);
  if (wid >= strips * chunks * nb) return;
  const int sx = wid % strips, cy = (wid / strips) % chunks, z = wid / (strips * chunks);
  rd_batch_off((size_t)z * fs, extH, extV, edge);
  const int x0 = sx * 32, x = x0 + lane;
  const int y0 = cy * ch, y1 = min(y0 + ch, ih);
  const int first = max(y0 - 5, 0);
  const bool okl = x - 32 >= 0, okc = x < iw, okr = x + 32 < iw;
  // positions of the x walk that do not exist: left of column 0 (all of the left word when x0 == 0), right of column iw - 1
  const unsigned negL = x0 == 0 ? 0xffffffffu : 0u;
  const unsigned invC = x0 + 32 <= iw ? 0u : (0xffffffffu << (iw - x0));
  const unsigned invR = x0 + 64 <= iw ? 0u : (x0 + 32 >= iw ? 0xffffffffu : (0xffffffffu << (iw - x0 - 32)));
  unsigned pl = 0, pc = 0, pr = 0;                      // previous row as bit words: columns x0-32.., x0.., x0+32..
  unsigned col = 0, colS = 0;                           // column histories: bit b = row (y - b) of column x / x + 1
  for (int y = first; y < y1 + 5; y++) {
    bool el = false, ec = false, er = false;
    if (y < ih) {
      const int8_t *row = edge + (size_t)y * iw + x;
      if (okl) el = row[-32] != 0;
      if (okc) ec = row[0] != 0;
      if (okr) er = row[32] != 0;
    }
    const unsigned cl = __ballot_sync(0xffffffffu, el), cc = __ballot_sync(0xffffffffu, ec), cr = __ballot_sync(0xffffffffu, er);
    // ---- x walk of row y - 1 (its perpendicular neighbour is row y; past the last row cc == 0 and the side rule is void)
    const int yh = y - 1;
    if (yh >= y0 && yh < y1) {
      const unsigned prevC = (pc << 1) | (pl >> 31), prevL = pl << 1;          // bit 0 of prevL belongs to a word never consulted
      unsigned sbC = exs_stop_back(pc, prevC, cc), sbL = exs_stop_back(pl, prevL, cl) | negL;
      if (x0 == 0) sbC &= ~1u;                                                   // no rule applies at position 0
      const unsigned nextC = (pc >> 1) | (pr << 31), nextR = pr >> 1;          // bit 31 of nextR likewise
      const unsigned faC = (~pc & nextC) | invC, faR = (~pr & nextR) | invR;    // forward stops when the origin is not an edge pixel
      const unsigned fbC = ~pc | invC, fbR = ~pr | invR;                        // ... and when it is
      const unsigned long long back = (((unsigned long long)sbC << 32) | sbL) >> (lane + 28);
      const bool oe = (pc >> lane) & 1u;
      const unsigned long long fwd = (((unsigned long long)(oe ? fbR : faR) << 32) | (oe ? fbC : faC)) >> lane;
      const unsigned tb = (unsigned)back & 31u, tf = (unsigned)fwd & 31u;
      const unsigned nl = __clz(tb) - 27, nr = tf ? __ffs(tf) - 1 : 5;
      if (okc) extH[(size_t)yh * iw + x] = (uint8_t)(nl | (nr << 4));
    }
    pl = cl; pc = cc; pr = cr;
    // ---- y walk of row y - 5: rows y - 10 .. y are bits 10 .. 0
    const unsigned long long both = ((unsigned long long)cr << 32) | cc;
    col = (col << 1) | (unsigned)ec;
    colS = (colS << 1) | ((unsigned)(both >> (lane + 1)) & 1u);
    const int yv = y - 5;
    if (yv >= y0 && yv < y1) {
      const bool hasSide = x < iw - 1;
      unsigned sb = exs_stop_back(col, col >> 1, hasSide ? colS : 0u);
      if (y < 31) sb |= 0xffffffffu << (y + 1);                                  // rows above the image
      if (y < 32) sb &= ~(1u << y);                                              // row 0
      const bool oe = (col >> 5) & 1u;
      unsigned fw = oe ? ~col : (~col & (col << 1));
      if (y >= ih) fw |= (2u << (y - ih)) - 1u;                                  // rows below the image
      const unsigned tb = (sb >> 5) & 31u, tf = (fw >> 1) & 31u;
      const unsigned nu = tb ? __ffs(tb) - 1 : 5, nd = __clz(tf) - 27;
      if (okc) extV[(size_t)yv * iw + x] = (uint8_t)(nu | (nd << 4));
    }
  }
}

// ---- the iteration kernel, four pixels per lane --------------------------------------------------------------------
// Same streaming scheme as kf_blb_stream on strips of 128 columns: a lane owns four consecutive pixels (one 128-bit
// load / store per row), so the shuffle scan of the row sums is paid once per four pixels and the apron of the x pass
// (four pixels each side, one 128-bit load by lane 0 and one by lane 1) shrinks from 25 % to 6 %.  The running sums of a
// row go to shared memory as P[i & 3][i >> 2] (i = position in the 136-pixel row), which keeps the lanes of a warp on
// different banks whatever their extents; the column sums live in a 16-row ring with one column per (lane, pixel).
// All sums are kept modulo 2^64; a window sum is a difference of two of them and always fits its 20-bit field.
#define S4_WARPS 4
// The y pass of row yv looks at the column sums behind rows yv - 4 .. yv + 5, so a ring of S4_RING = 10 rows is enough
// (10 KB per warp instead of 16 with a power-of-two depth: 16 warps per SM instead of 12, and room for other kernels'
// CTAs beside them).  Slots are addressed relative to the newest row, `top` = slot of row yv + 5: row yv + 5 - k lives
// in slot top - k (+ S4_RING if negative).
#define S4_RING 10
template <int NPX> struct S4Smem {
  enum { AC = 4 / NPX, PP = 32 + 2 * AC + 2 };                    // apron columns per side; pitch of a plane of running sums
  blb_w P[2][NPX][PP];
  blb_w R[S4_RING][NPX][32];
};
__device__ __forceinline__ int s4_slot(int top, int k) { const int t = top - k; return t < 0 ? t + S4_RING : t; }
template <int NPX> struct S4Vec;
template <> struct S4Vec<4> { typedef uint4 T; };
template <> struct S4Vec<2> { typedef uint2 T; };
__device__ __forceinline__ void s4_get(const uint4 &v, uint32_t (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ void s4_get(const uint2 &v, uint32_t (&o)[2]) { o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ uint4 s4_make(const uint32_t (&r)[4]) { return make_uint4(r[0], r[1], r[2], r[3]); }
__device__ __forceinline__ uint2 s4_make(const uint32_t (&r)[2]) { return make_uint2(r[0], r[1]); }
// floor(c / w) = umulhi(2c, ceil(2^31 / w)) for c < 2^16, 1 <= w <= 10
__constant__ const unsigned S4_M31[16] = {0u, 0x80000000u, 0x40000000u, 0x2aaaaaabu, 0x20000000u, 0x1999999au, 0x15555556u, 0x12492493u, 0x10000000u, 0x0e38e38fu, 0x0ccccccdu, 0u, 0u, 0u, 0u, 0u};
__device__ __forceinline__ void s4_fields2(blb_w s, unsigned &f0, unsigned &f1, unsigned &f2) {      // the three fields, doubled
  const unsigned lo = (unsigned)s, hi = (unsigned)(s >> 32);
  f0 = (lo << 1) & 0x1ffffeu; f1 = __funnelshift_r(lo, hi, 19) & 0x1ffffeu; f2 = (hi >> 7) & 0x1ffffeu;
}
__device__ __forceinline__ blb_w s4_mean_spread(blb_w s, unsigned m) {
  unsigned f0, f1, f2;
  s4_fields2(s, f0, f1, f2);
  const unsigned q0 = __umulhi(f0, m), q1 = __umulhi(f1, m), q2 = __umulhi(f2, m);
  return ((blb_w)(q2 << 8) << 32) | (blb_w)(q0 | (q1 << 20));
}
__device__ __forceinline__ uint32_t s4_mean_packed(blb_w s, unsigned m) {
  unsigned f0, f1, f2;
  s4_fields2(s, f0, f1, f2);
  return (__umulhi(f2, m) << 22) | (__umulhi(f1, m) << 12) | __umulhi(f0, m);
}
// y pass for row yv of the lane's NPX columns: the sums behind rows yv - 4 .. yv + 5 are in the ring; top = slot of row yv + 5
template <int NPX>
__device__ __forceinline__ void s4_vrow(uint32_t (&res)[NPX], const blb_w (*R)[NPX][32], const unsigned *rcp, unsigned ev, int top, int lane) {
  const int sa = s4_slot(top, 4), sd = s4_slot(top, 5);            // rows yv + 1, yv
#pragma unroll
  for (int k = 0; k < NPX; k++) {
    const unsigned ek = (ev >> (8 * k)) & 255u, nu = ek & 15u, nd = ek >> 4;
    const blb_w qa = R[sa][k][lane], qd = R[sd][k][lane];
    const blb_w vs = (qa - R[s4_slot(top, 4 + (int)nu)][k][lane]) + (R[s4_slot(top, 5 - (int)nd)][k][lane] - qd);
    const unsigned ws = nu + nd;
    res[k] = ws == 0 ? blb_pack(qa - qd) : s4_mean_packed(vs, rcp[ws]);
  }
}
template <int NPX>
__global__ void __launch_bounds__(S4_WARPS * 32) kf_blb_stream4(uint32_t *out, const uint32_t *in, const uint8_t *extH, const uint8_t *extV, int iw, int ih, int nb,
                                                                int strips, int chunks, int ch, size_t fs) {
  typedef S4Smem<NPX> Smem;
  typedef typename S4Vec<NPX>::T Vec;
  constexpr int AC = Smem::AC, SW = 32 * NPX;                     // strip width
  extern __shared__ __align__(16) unsigned char s4_raw[];
  __shared__ unsigned rcp[16];
  if (threadIdx.x < 16) rcp[threadIdx.x] = S4_M31[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int wid = blockIdx.x * S4_WARPS + wp;
  if (wid >= strips * chunks * nb) return;
  Smem &sm = ((Smem *)s4_raw)[wp];
  const int sx = wid % strips, cy = (wid / strips) % chunks, z = wid / (strips * chunks);
  rd_batch_off((size_t)z * fs, out, in, extH, extV);
  const int x0 = sx * SW, x = x0 + NPX * lane;
  const bool okx = x < iw;
  const bool okap = (lane == 0 && sx > 0) || (lane == 1 && x0 + SW < iw);
  const int y0 = cy * ch, y1 = min(y0 + ch, ih);
  const int first = max(y0 - BLB, 0), last = min(y1 + BLB, ih);
  const Vec *inv = (const Vec *)in;
  const uint4 *in4 = (const uint4 *)in;
  Vec *outv = (Vec *)out;
  const int qv = iw / NPX, cx = x / NPX;                          // row pitch and column in vector units
  const int q4 = iw >> 2, cap = (lane == 0 ? x0 - 4 : x0 + SW) >> 2;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  Vec zerov;
  memset(&zerov, 0, sizeof(zerov));
  blb_w Q[NPX];
  int slot = first % S4_RING;                                     // slot of the row whose column sums were stored last
#pragma unroll
  for (int k = 0; k < NPX; k++) { Q[k] = 0; sm.R[slot][k][lane] = 0; }
  auto load_ext = [&](const uint8_t *e, int row) -> unsigned {
    if (NPX == 4) return ((const uint32_t *)e)[(size_t)row * qv + cx];
    return ((const uint16_t *)e)[(size_t)row * qv + cx];
  };
  // two rows of look-ahead in registers
  Vec v0 = zerov, v1 = zerov;
  uint4 a0 = zero4, a1 = zero4;
  unsigned e0 = 0, e1 = 0;
  {
    if (okx) { v0 = inv[(size_t)first * qv + cx]; e0 = load_ext(extH, first); }
    if (okap) a0 = in4[(size_t)first * q4 + cap];
    if (first + 1 < last) {
      if (okx) { v1 = inv[(size_t)(first + 1) * qv + cx]; e1 = load_ext(extH, first + 1); }
      if (okap) a1 = in4[(size_t)(first + 1) * q4 + cap];
    }
  }
  for (int y = first; y < last; y++) {
    const Vec v = v0;
    const uint4 ap = a0;
    const unsigned e = e0;
    v0 = v1; a0 = a1; e0 = e1;
    v1 = zerov; a1 = zero4; e1 = 0;
    if (y + 2 < last) {
      if (okx) { v1 = inv[(size_t)(y + 2) * qv + cx]; e1 = load_ext(extH, y + 2); }
      if (okap) a1 = in4[(size_t)(y + 2) * q4 + cap];
    }
    const int yv = y - BLB;
    const bool doV = yv >= y0 && okx;
    unsigned ev = 0;
    if (doV) ev = load_ext(extV, yv);
    // ---- x pass: running sums of the row (4 apron pixels, 32 * NPX pixels, 4 apron pixels)
    uint32_t pv[NPX];
    s4_get(v, pv);
    blb_w w[NPX], pc[NPX + 1];
#pragma unroll
    for (int k = 0; k < NPX; k++) w[k] = blb_spread(pv[k]);
    blb_w tot = w[0];
#pragma unroll
    for (int k = 1; k < NPX; k++) tot += w[k];
    const blb_w aw0 = blb_spread(ap.x), aw1 = blb_spread(ap.y), aw2 = blb_spread(ap.z), aw3 = blb_spread(ap.w);
    const blb_w as1 = aw0 + aw1, as2 = as1 + aw2, as3 = as2 + aw3;
    const blb_w S = blb_scan_up(tot, lane, 5);
    const blb_w T = __shfl_sync(0xffffffffu, S, 31), LA = __shfl_sync(0xffffffffu, as3, 0);
    pc[0] = LA + S - tot;                                           // running sum in front of the lane's first pixel
#pragma unroll
    for (int k = 0; k < NPX; k++) pc[k + 1] = pc[k] + w[k];
    blb_w (*P)[Smem::PP] = sm.P[y & 1];
#pragma unroll
    for (int k = 0; k < NPX; k++) P[k][lane + AC] = pc[k];
    // position i of the row lives in P[i % NPX][i / NPX]
#define S4_P(i) P[(i) % NPX][(i) / NPX]
    if (lane == 0) { S4_P(0) = 0; S4_P(1) = aw0; S4_P(2) = as1; S4_P(3) = as2; }
    if (lane == 1) { const blb_w b = LA + T; S4_P(4 + SW) = b; S4_P(5 + SW) = b + aw0; S4_P(6 + SW) = b + as1; S4_P(7 + SW) = b + as2; S4_P(8 + SW) = b + as3; }
#undef S4_P
    __syncwarp();
    blb_w h[NPX];
#pragma unroll
    for (int k = 0; k < NPX; k++) {
      const unsigned ek = (e >> (8 * k)) & 255u, nl = ek & 15u, nr = ek >> 4;
      const unsigned j1 = k + 5 - nl, j2 = k + 4 + nr;            // positions relative to the lane's first column
      const blb_w hs = (pc[k + 1] - P[j1 % NPX][lane + j1 / NPX]) + (P[j2 % NPX][lane + j2 / NPX] - pc[k]);
      const unsigned ws = nl + nr;
      h[k] = ws == 0 ? w[k] : s4_mean_spread(hs, rcp[ws]);
    }
    // ---- y pass: column running sums in the ring; row yv = y - 4 is complete once the sums behind row y are known
    slot = slot + 1 == S4_RING ? 0 : slot + 1;                     // row y + 1 = yv + 5
#pragma unroll
    for (int k = 0; k < NPX; k++) { Q[k] += h[k]; sm.R[slot][k][lane] = Q[k]; }
    if (doV) {
      uint32_t res[NPX];
      s4_vrow<NPX>(res, sm.R, rcp, ev, slot, lane);
      outv[(size_t)yv * qv + cx] = s4_make(res);
    }
  }
  // rows whose downward walk is cut by the bottom of the image rather than by the strip
  if (okx) {
    for (int yv = max(y0, last - BLB); yv < y1; yv++) {
      uint32_t res[NPX];
      s4_vrow<NPX>(res, sm.R, rcp, load_ext(extV, yv), (yv + 5) % S4_RING, lane);
      outv[(size_t)yv * qv + cx] = s4_make(res);
    }
  }
}

static int rd_sm_count() {
  static int sms = 0;
  if (!sms) { int dev = 0; RD_CUDA(cudaGetDevice(&dev)); RD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)); }
  return sms;
}
// strip height for a streaming kernel: enough strips to fill `warps_per_sm` warp slots of every SM in one wave, but not
// shorter than `min_ch` rows (each strip re-reads its apron rows) nor longer than `max_ch`
static int rd_strip_height(int strips, int nb, int ih, int warps_per_sm, int min_ch, int max_ch) {
  int chunks = (rd_sm_count() * warps_per_sm) / (strips * nb);
  if (chunks < 1) chunks = 1;
  int ch = rd_cdiv(ih, chunks);
  if (ch < min_ch) ch = min_ch;
  if (ch > max_ch) ch = max_ch;
  return ch;
}

// steps 13 of genGPUTask: src -> 10 iterations -> dst, ping-ponging through `pong`; ext: 2*iw*ih bytes of scratch
void rd_blblur_run(uint32_t *dst, uint32_t *pong, const uint32_t *src, const int8_t *edge, uint8_t *ext, int iters, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  uint8_t *extH = ext, *extV = ext + (size_t)iw * ih;
  {
    // (a single frame is cut into short strips: latency matters more than the re-read apron rows when the machine is empty;
    // measured at nb = 1, 1280x720: strips of 13 / 8 rows -> 27.2 / 23.6 us per iteration)
    const int strips = rd_cdiv(iw, 32), ch = rd_strip_height(strips, nb, ih, 32, nb >= 8 ? 40 : 8 + 4 * nb, 1 << 20), chunks = rd_cdiv(ih, ch);
    RD_LAUNCH(kf_blb_extents_s, rd_cdiv(strips * chunks * nb, EXS_WARPS), EXS_WARPS * 32, 0, s, extH, extV, edge, iw, ih, nb, strips, chunks, ch, fs);
  }
  const uint32_t *cur = src;
  if ((iw & 3) == 0) {
    // Four pixels per lane: 12.3 KB of shared memory per warp, 16 warps per SM.  (Measured: the two-pixel instantiation
    // - twice the warps per SM - is 10 % slower, 8.0 vs 7.2 us per 1280x720 iteration; the kernel is bound by instruction issue at
    // ~170 instructions per pixel, not by latency, and the scan and apron cost more per pixel with narrower strips.)
    constexpr int npx = 4;
    static bool attr = false;
    const size_t smem = S4_WARPS * sizeof(S4Smem<npx>);
    if (!attr) { RD_CUDA(cudaFuncSetAttribute(kf_blb_stream4<npx>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    static const int minch_env = getenv("RD_BLB_MINCH") ? atoi(getenv("RD_BLB_MINCH")) : 0;     // (experiments)
    const int strips = rd_cdiv(iw, 32 * npx), ch = rd_strip_height(strips, nb, ih, 16, minch_env > 0 && nb < 8 ? minch_env : (nb >= 8 ? 48 : 4 + 4 * nb), 1 << 20), chunks = rd_cdiv(ih, ch);
    const int blocks = rd_cdiv(strips * chunks * nb, S4_WARPS);
    for (int i = 0; i < iters; i++) {
      uint32_t *o = ((iters - i) & 1) ? dst : pong;     // the last iteration lands in dst
      RD_LAUNCH(kf_blb_stream4<npx>, blocks, S4_WARPS * 32, smem, s, o, cur, extH, extV, iw, ih, nb, strips, chunks, ch, fs);
      cur = o;
    }
    return;
  }
  // widths that are not a multiple of four: one pixel per lane.  Strip height: as many strips as fit the machine in one wave
  // (4 CTAs of 8 warps per SM at 58 registers); clamped so that strips stay worth their apron
  const int strips = rd_cdiv(iw, 32), ch = rd_strip_height(strips, nb, ih, 4 * SB_WARPS, SB_CH_MIN, SB_CH_MAX), chunks = rd_cdiv(ih, ch);
  const int blocks = rd_cdiv(strips * chunks * nb, SB_WARPS);
  for (int i = 0; i < iters; i++) {
    uint32_t *o = ((iters - i) & 1) ? dst : pong;       // the last iteration lands in dst
    RD_LAUNCH(kf_blb_stream, blocks, SB_WARPS * 32, 0, s, o, cur, extH, extV, iw, ih, nb, strips, chunks, ch, fs);
    cur = o;
  }
}

// =============================================================================================== calcSize
#define CS_SLOTS 64
__global__ void __launch_bounds__(256) kf_calcSize(int *out, const int *label, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, label);
  __shared__ int keys[CS_SLOTS];
  __shared__ int cnts[CS_SLOTS];
  if (threadIdx.x < CS_SLOTS) { keys[threadIdx.x] = -1; cnts[threadIdx.x] = 0; }
  __syncthreads();
  const int lx = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int x = blockIdx.x * 32 + lx;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int y = blockIdx.y * 32 + wy + k * 8;
    const bool ok = x < iw && y < ih;
    const int l = ok ? label[(size_t)y * iw + x] : -1;
    const unsigned act = __ballot_sync(0xffffffffu, l != -1);
    if (l == -1) continue;
    const unsigned peers = __match_any_sync(act, l);
    if ((int)(__ffs(peers) - 1) != lx) continue;
    const int c = __popc(peers);
    unsigned h = ((unsigned)l * 2654435761u) >> 26;           // 6-bit hash
    bool done = false;
    for (int probe = 0; probe < 8 && !done; probe++, h = (h + 1) & (CS_SLOTS - 1)) {
      const int prev = atomicCAS(&keys[h], -1, l);
      if (prev == -1 || prev == l) { atomicAdd(&cnts[h], c); done = true; }
    }
    if (!done) atomicAdd(out + l, c);                         // crowded tile: straight to global
  }
  __syncthreads();
  if (threadIdx.x < CS_SLOTS && keys[threadIdx.x] != -1) atomicAdd(out + keys[threadIdx.x], cnts[threadIdx.x]);
}
void rd_calcSize_run(int *out, const int *label, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  RD_LAUNCH(kf_calcSize, dim3(rd_cdiv(iw, 32), rd_cdiv(ih, 32), nb), 256, 0, s, out, label, iw, ih, fs);
}

// =============================================================================================== recursive Gaussian x3
// Horizontal: one thread per (row, direction) runs the L, a and b chains of that row straight from the packed Lab
// plane (the unpack of oclimgutil.cl:333 is three ALU ops) and stores the raw causal / anti-causal responses.
// Each lane walks its own row with 128-bit loads / stores; the L1 keeps the other half of each sector for the next step.
// unpack one channel of a packed Lab word (oclimgutil.cl:36-39)
template <int CH> __device__ __forceinline__ float unpack_ch(uint32_t plab) {
  const int sh = CH == 0 ? 0 : (CH == 1 ? 12 : 22);
  const unsigned mask = CH == 0 ? 4095u : 1023u;
  const float sc = CH == 0 ? 1.0f / 4096 : 1.0f / 1024, of = CH == 0 ? 0.5f / 4096 : 0.5f / 1024;
  return __fadd_rn(__fmul_rn((float)(int)((plab >> sh) & mask), sc), of);
}
// The recurrence (oclimgutil.cl:549-558) eight steps at a time.  The taps are the previous seven inputs and outputs; with
// the loop fully unrolled they are plain registers named at compile time, so no tap is ever moved (the straightforward
// form shifts 14 registers per step, which doubled the instruction count).  pi / po: the last eight inputs / outputs,
// index 7 = newest.  Order of operations per step is that of rd_iir_taps::step.
struct Chain8 {
  float pi[8], po[8];
  __device__ __forceinline__ Chain8() {
#pragma unroll
    for (int i = 0; i < 8; i++) pi[i] = po[i] = 0.0f;
  }
  __device__ __forceinline__ float step1(float in, const float *c) {          // warm-up / ragged tail: one step, taps shifted
    float d = __fmul_rn(in, c[0]);
    float a = __fmul_rn(c[1], pi[7]);
#pragma unroll
    for (int j = 2; j <= 7; j++) a = __fadd_rn(a, __fmul_rn(c[j], pi[8 - j]));
    d = __fadd_rn(d, a);
    float b = __fmul_rn(c[8], po[7]);
#pragma unroll
    for (int j = 1; j <= 6; j++) b = __fadd_rn(b, __fmul_rn(c[8 + j], po[7 - j]));
    d = __fadd_rn(d, b);
#pragma unroll
    for (int i = 0; i < 7; i++) { pi[i] = pi[i + 1]; po[i] = po[i + 1]; }
    pi[7] = in; po[7] = d;
    return d;
  }
  __device__ __forceinline__ void step8(const float (&in)[8], float (&out)[8], const float *c) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      // X(j): input j steps into this block (j >= 0) or from the previous block (j < 0); O(j) likewise for outputs
#define X_(j) ((j) >= 0 ? in[(j) < 0 ? 0 : (j)] : pi[(j) < 0 ? 8 + (j) : 0])
#define O_(j) ((j) >= 0 ? out[(j) < 0 ? 0 : (j)] : po[(j) < 0 ? 8 + (j) : 0])
      float d = __fmul_rn(in[k], c[0]);
      float a = __fmul_rn(c[1], X_(k - 1));
#pragma unroll
      for (int j = 2; j <= 7; j++) a = __fadd_rn(a, __fmul_rn(c[j], X_(k - j)));
      d = __fadd_rn(d, a);
      float b = __fmul_rn(c[8], O_(k - 1));
#pragma unroll
      for (int j = 1; j <= 6; j++) b = __fadd_rn(b, __fmul_rn(c[8 + j], O_(k - 1 - j)));
      d = __fadd_rn(d, b);
      out[k] = d;
#undef X_
#undef O_
    }
#pragma unroll
    for (int i = 0; i < 8; i++) { pi[i] = in[i]; po[i] = out[i]; }
  }
};

// Horizontal pass: thread t -> channel t % 3, direction (t / 3) & 1, row t / 6 (the six chains of a row sit in
// neighbouring lanes).  Each lane streams its row with 128-bit loads (next eight pixels prefetched) and 128-bit stores.
template <int CH, int DIR>
__device__ __forceinline__ void iir_h_chain(float *q, const uint32_t *row, const float *c, int warm, int iw) {
  Chain8 st;
  if (DIR == 0) for (int x = -warm; x < 0; x++) st.step1(unpack_ch<CH>(row[rd_mirror1(x, iw)]), c);
  else for (int x = iw + warm; x >= iw; x--) st.step1(unpack_ch<CH>(row[rd_mirror1(x, iw)]), c);
  if ((iw & 7) == 0) {
    const int nblk = iw >> 3;
    // block b covers pixels [8b, 8b+8) walking forward, [iw-8-8b, iw-8b) walking backward
    const uint4 *src = (const uint4 *)row;
    // two blocks of look-ahead: a step8 is ~250 dependent instructions, an HBM/L2 round trip for 32 scattered sectors is longer
    uint4 v0 = src[DIR == 0 ? 0 : 2 * nblk - 2], v1 = src[DIR == 0 ? 1 : 2 * nblk - 1];
    uint4 u0 = v0, u1 = v1;
    if (nblk > 1) { u0 = src[DIR == 0 ? 2 : 2 * nblk - 4]; u1 = src[DIR == 0 ? 3 : 2 * nblk - 3]; }
    for (int b = 0; b < nblk; b++) {
      const int base = DIR == 0 ? 2 * b : 2 * (nblk - 1 - b);
      const uint4 c0 = v0, c1 = v1;
      v0 = u0; v1 = u1;
      if (b + 2 < nblk) { const int nb2 = DIR == 0 ? base + 4 : base - 4; u0 = src[nb2]; u1 = src[nb2 + 1]; }
      const uint32_t w[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
      float in[8], out[8];
#pragma unroll
      for (int k = 0; k < 8; k++) in[k] = unpack_ch<CH>(w[DIR == 0 ? k : 7 - k]);
      st.step8(in, out, c);
      float4 *dst = (float4 *)q + base;
      if (DIR == 0) { dst[0] = make_float4(out[0], out[1], out[2], out[3]); dst[1] = make_float4(out[4], out[5], out[6], out[7]); }
      else { dst[0] = make_float4(out[7], out[6], out[5], out[4]); dst[1] = make_float4(out[3], out[2], out[1], out[0]); }
    }
  } else {
    if (DIR == 0) for (int x = 0; x < iw; x++) q[x] = st.step1(unpack_ch<CH>(row[x]), c);
    else for (int x = iw - 1; x >= 0; x--) q[x] = st.step1(unpack_ch<CH>(row[x]), c);
  }
}
// CTA = 6 warps over 32 rows: warp w runs channel w / 2 in direction w % 2 (no divergence inside a warp), lane = row.
__global__ void __launch_bounds__(192) kf_iir_h3(float *f0, float *f1, float *f2, float *b0, float *b1, float *b2, const uint32_t *plab, int r, int iw, int ih, size_t fs) {
  rd_batch_y(fs, f0, f1, f2, b0, b1, b2, plab);
  const int w = threadIdx.x >> 5, y = blockIdx.x * 32 + (threadIdx.x & 31);
  if (y >= ih) return;
  float c[15];
#pragma unroll
  for (int i = 0; i < 15; i++) c[i] = RD_IIRCOEF[r][i];
  const uint32_t *row = plab + (size_t)y * iw;
  const size_t ro = (size_t)y * iw;
  const int warm = r + 1 + 8;
  switch (w) {
    case 0: iir_h_chain<0, 0>(f0 + ro, row, c, warm, iw); break;
    case 1: iir_h_chain<0, 1>(b0 + ro, row, c, warm, iw); break;
    case 2: iir_h_chain<1, 0>(f1 + ro, row, c, warm, iw); break;
    case 3: iir_h_chain<1, 1>(b1 + ro, row, c, warm, iw); break;
    case 4: iir_h_chain<2, 0>(f2 + ro, row, c, warm, iw); break;
    default: iir_h_chain<2, 1>(b2 + ro, row, c, warm, iw); break;
  }
}
// pass1 (oclimgutil.cl:580) for the three channels: h = bwd + fwd - in*coef0, in place of the forward planes; 4 pixels per thread
__device__ __forceinline__ float iir_comb(float b, float f, float in, float c0) { return __fsub_rn(__fadd_rn(b, f), __fmul_rn(in, c0)); }
__global__ void kf_iir_mid3(float *f0, float *f1, float *f2, const float *b0, const float *b1, const float *b2, const uint32_t *plab, int r, int n, size_t fs) {
  rd_batch_y(fs, f0, f1, f2, b0, b1, b2, plab);
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float c0 = RD_IIRCOEF[r][0];
  if (i + 3 < n) {
    const uint4 p = *(const uint4 *)(plab + i);
    const uint32_t pw[4] = {p.x, p.y, p.z, p.w};
    float4 F0 = *(float4 *)(f0 + i), F1 = *(float4 *)(f1 + i), F2 = *(float4 *)(f2 + i);
    const float4 B0 = *(const float4 *)(b0 + i), B1 = *(const float4 *)(b1 + i), B2 = *(const float4 *)(b2 + i);
    float *pf0 = (float *)&F0, *pf1 = (float *)&F1, *pf2 = (float *)&F2;
    const float *pb0 = (const float *)&B0, *pb1 = (const float *)&B1, *pb2 = (const float *)&B2;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float l, a, b;
      rd_unpacklab(pw[k], l, a, b);
      pf0[k] = iir_comb(pb0[k], pf0[k], l, c0); pf1[k] = iir_comb(pb1[k], pf1[k], a, c0); pf2[k] = iir_comb(pb2[k], pf2[k], b, c0);
    }
    *(float4 *)(f0 + i) = F0; *(float4 *)(f1 + i) = F1; *(float4 *)(f2 + i) = F2;
  } else {
    for (int k = i; k < n; k++) {
      float l, a, b;
      rd_unpacklab(plab[k], l, a, b);
      f0[k] = iir_comb(b0[k], f0[k], l, c0); f1[k] = iir_comb(b1[k], f1[k], a, c0); f2[k] = iir_comb(b2[k], f2[k], b, c0);
    }
  }
}
// Vertical: one thread per (column, channel, direction), coalesced across the warp; the rows of the next eight steps are
// loaded while the current eight are computed.
__global__ void __launch_bounds__(128) kf_iir_v3(float *vf0, float *vf1, float *vf2, float *vb0, float *vb1, float *vb2,
                                                 const float *h0, const float *h1, const float *h2, int r, int iw, int ih, size_t fs) {
  rd_batch_z(fs, vf0, vf1, vf2, vb0, vb1, vb2, h0, h1, h2);
  const int x = blockIdx.x * blockDim.x + threadIdx.x, ch = blockIdx.y >> 1, dir = blockIdx.y & 1;
  if (x >= iw) return;
  float c[15];
#pragma unroll
  for (int i = 0; i < 15; i++) c[i] = RD_IIRCOEF[r][i];
  const float *__restrict__ h = (ch == 0 ? h0 : (ch == 1 ? h1 : h2)) + x;
  float *__restrict__ o = (dir == 0 ? (ch == 0 ? vf0 : (ch == 1 ? vf1 : vf2)) : (ch == 0 ? vb0 : (ch == 1 ? vb1 : vb2))) + x;
  Chain8 st;
  const int warm = r + 1 + 8;
  const int sgn = dir == 0 ? 1 : -1;
  if (dir == 0) for (int y = -warm; y < 0; y++) st.step1(h[(size_t)rd_mirror1(y, ih) * iw], c);
  else for (int y = ih + warm; y >= ih; y--) st.step1(h[(size_t)rd_mirror1(y, ih) * iw], c);
  const int y0 = dir == 0 ? 0 : ih - 1;                          // first real row; row k of the walk is y0 + k*sgn
  const int nblk = ih >> 3;
  // running pointers: row k of the walk is hp[k * step] (64-bit index arithmetic per element cost a quarter of the kernel)
  const ptrdiff_t step = (ptrdiff_t)sgn * iw;
  const float *hp = h + (size_t)y0 * iw;
  float *op = o + (size_t)y0 * iw;
  float cur[8], nxt[8];
#pragma unroll
  for (int j = 0; j < 8; j++) cur[j] = j < ih ? hp[j * step] : 0.0f;
  for (int b = 0; b < nblk; b++) {
    const int k0 = b * 8;
    if (k0 + 16 <= ih) {
#pragma unroll
      for (int j = 0; j < 8; j++) nxt[j] = hp[(8 + j) * step];
    } else {
#pragma unroll
      for (int j = 0; j < 8; j++) nxt[j] = k0 + 8 + j < ih ? hp[(8 + j) * step] : 0.0f;
    }
    float out[8];
    st.step8(cur, out, c);
#pragma unroll
    for (int j = 0; j < 8; j++) op[j * step] = out[j];
#pragma unroll
    for (int j = 0; j < 8; j++) cur[j] = nxt[j];
    hp += 8 * step; op += 8 * step;
  }
  for (int k = nblk * 8, j = 0; k < ih; k++, j++) {               // ragged tail (ih % 8 rows): cur[] holds them
    float v = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; q++) if (q == j) v = cur[q];
    op[j * step] = st.step1(v, c);
  }
}
// pass3 (oclimgutil.cl:629) for the three channels + pack_plab (oclimgutil.cl:325): blurred L plane and blurred packed Lab;
// outL / outA / outB alias h0 / h1 / h2 (each thread reads its own four elements before writing them)
__global__ void kf_iir_fin3(float *outL, float *outA, float *outB, uint32_t *outPlab, const float *vf0, const float *vf1, const float *vf2, const float *vb0,
                            const float *vb1, const float *vb2, int r, int n, size_t fs) {
  rd_batch_y(fs, outL, outA, outB, outPlab, vf0, vf1, vf2, vb0, vb1, vb2);
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float c0 = RD_IIRCOEF[r][0];
  if (i + 3 < n) {
    float4 H0 = *(float4 *)(outL + i), H1 = *(float4 *)(outA + i), H2 = *(float4 *)(outB + i);
    const float4 F0 = *(const float4 *)(vf0 + i), F1 = *(const float4 *)(vf1 + i), F2 = *(const float4 *)(vf2 + i);
    const float4 B0 = *(const float4 *)(vb0 + i), B1 = *(const float4 *)(vb1 + i), B2 = *(const float4 *)(vb2 + i);
    float *h0 = (float *)&H0, *h1 = (float *)&H1, *h2 = (float *)&H2;
    const float *pf0 = (const float *)&F0, *pf1 = (const float *)&F1, *pf2 = (const float *)&F2, *pb0 = (const float *)&B0, *pb1 = (const float *)&B1, *pb2 = (const float *)&B2;
    uint4 P;
    uint32_t *pp = (uint32_t *)&P;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      h0[k] = iir_comb(pb0[k], pf0[k], h0[k], c0); h1[k] = iir_comb(pb1[k], pf1[k], h1[k], c0); h2[k] = iir_comb(pb2[k], pf2[k], h2[k], c0);
      pp[k] = rd_packlab(h0[k], h1[k], h2[k]);
    }
    *(float4 *)(outL + i) = H0; *(float4 *)(outA + i) = H1; *(float4 *)(outB + i) = H2;
    *(uint4 *)(outPlab + i) = P;
  } else {
    for (int k = i; k < n; k++) {
      const float l = iir_comb(vb0[k], vf0[k], outL[k], c0), a = iir_comb(vb1[k], vf1[k], outA[k], c0), b = iir_comb(vb2[k], vf2[k], outB[k], c0);
      outL[k] = l; outA[k] = a; outB[k] = b;
      outPlab[k] = rd_packlab(l, a, b);
    }
  }
}

// =============================================================================================== BGR8 -> packed Lab
// bgr2plab (oclimgutil.cl:256) four pixels per thread: three 32-bit loads bring 12 bytes, one 128-bit store leaves.
__global__ void __launch_bounds__(256) kf_bgr2plab4(uint32_t *out, const uint8_t *in, size_t in_fs, int iw, int ih, int ws, size_t fs) {
  rd_batch_z(fs, out);
  rd_batch_z(in_fs, in);
  const int x4 = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x4 * 4 >= iw || y >= ih) return;
  const uint32_t *p = (const uint32_t *)(in + (size_t)y * ws) + x4 * 3;
  const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
  uint4 r;
  r.x = rd_srgb2plab(w0 & 255, (w0 >> 8) & 255, (w0 >> 16) & 255, RD_S2L, RD_CFUNC, RD_CFUNC2);
  r.y = rd_srgb2plab(w0 >> 24, w1 & 255, (w1 >> 8) & 255, RD_S2L, RD_CFUNC, RD_CFUNC2);
  r.z = rd_srgb2plab((w1 >> 16) & 255, w1 >> 24, w2 & 255, RD_S2L, RD_CFUNC, RD_CFUNC2);
  r.w = rd_srgb2plab((w2 >> 8) & 255, (w2 >> 16) & 255, w2 >> 24, RD_S2L, RD_CFUNC, RD_CFUNC2);
  *(uint4 *)(out + (size_t)y * iw + x4 * 4) = r;
}
__global__ void kf_bgr2plab1(uint32_t *out, const uint8_t *in, size_t in_fs, int iw, int ih, int ws, size_t fs) {
  rd_batch_z(fs, out);
  rd_batch_z(in_fs, in);
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  const uint8_t *p = in + (size_t)y * ws + x * 3;
  out[(size_t)y * iw + x] = rd_srgb2plab(p[0], p[1], p[2], RD_S2L, RD_CFUNC, RD_CFUNC2);
}
// NV12 input (decoded-video frames straight into Stage A): Y plane (row stride ys) followed by the interleaved
// half-resolution UV plane (same stride).  YUV -> BGR is the integer BT.601 limited-range conversion of OpenCV's
// cvtColor(COLOR_YUV2BGR_NV12) - what cv::VideoCapture hands the reference's programs (vidrect.cpp:160-166) - fused with bgr2plab:
// the BGR frame never exists.  Pinned bit-exactly to OpenCV by tests/golden/nv12_golden.json.
__device__ __forceinline__ uint32_t nv12_px(int Y, int u, int v) {                 // u, v already minus 128
  const int y = max(0, Y - 16) * 1220542 + (1 << 19);
  const int r = (y + 1673527 * v) >> 20, g = (y - 852492 * v - 409993 * u) >> 20, b = (y + 2116026 * u) >> 20;
  return rd_srgb2plab(min(max(b, 0), 255), min(max(g, 0), 255), min(max(r, 0), 255), RD_S2L, RD_CFUNC, RD_CFUNC2);
}
__global__ void __launch_bounds__(256) kf_nv12_plab4(uint32_t *out, const uint8_t *in, size_t in_fs, int iw, int ih, int ys, size_t fs) {
  rd_batch_z(fs, out);
  rd_batch_z(in_fs, in);
  const int x4 = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x4 * 4 >= iw || y >= ih) return;
  const uint32_t yy = *(const uint32_t *)(in + (size_t)y * ys + x4 * 4);
  const uint32_t uv = *(const uint32_t *)(in + (size_t)ih * ys + (size_t)(y >> 1) * ys + x4 * 4);     // U0 V0 U1 V1
  const int u0 = (int)(uv & 255) - 128, v0 = (int)((uv >> 8) & 255) - 128, u1 = (int)((uv >> 16) & 255) - 128, v1 = (int)(uv >> 24) - 128;
  uint4 r;
  r.x = nv12_px(yy & 255, u0, v0); r.y = nv12_px((yy >> 8) & 255, u0, v0);
  r.z = nv12_px((yy >> 16) & 255, u1, v1); r.w = nv12_px(yy >> 24, u1, v1);
  *(uint4 *)(out + (size_t)y * iw + x4 * 4) = r;
}
__global__ void kf_nv12_plab1(uint32_t *out, const uint8_t *in, size_t in_fs, int iw, int ih, int ys, size_t fs) {
  rd_batch_z(fs, out);
  rd_batch_z(in_fs, in);
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  const uint8_t *uv = in + (size_t)ih * ys + (size_t)(y >> 1) * ys + (x & ~1);
  out[(size_t)y * iw + x] = nv12_px(in[(size_t)y * ys + x], (int)uv[0] - 128, (int)uv[1] - 128);
}
// ws > 0: BGR8 rows of ws bytes; ws < 0: NV12 with a row stride of -ws bytes
void rd_bgr2plab_run(uint32_t *out, const uint8_t *in, size_t in_fs, int iw, int ih, int ws, int nb, size_t fs, cudaStream_t s) {
  const dim3 b(32, 8);
  if (ws < 0) {
    const int ys = -ws;
    if ((iw & 3) == 0 && (ys & 3) == 0 && (in_fs & 3) == 0 && ((uintptr_t)in & 3) == 0)
      RD_LAUNCH(kf_nv12_plab4, dim3(rd_cdiv(iw / 4, 32), rd_cdiv(ih, 8), nb), b, 0, s, out, in, in_fs, iw, ih, ys, fs);
    else
      RD_LAUNCH(kf_nv12_plab1, dim3(rd_cdiv(iw, 32), rd_cdiv(ih, 8), nb), b, 0, s, out, in, in_fs, iw, ih, ys, fs);
    return;
  }
  if ((iw & 3) == 0 && (ws & 3) == 0 && (in_fs & 3) == 0 && ((uintptr_t)in & 3) == 0)
    RD_LAUNCH(kf_bgr2plab4, dim3(rd_cdiv(iw / 4, 32), rd_cdiv(ih, 8), nb), b, 0, s, out, in, in_fs, iw, ih, ws, fs);
  else
    RD_LAUNCH(kf_bgr2plab1, dim3(rd_cdiv(iw, 32), rd_cdiv(ih, 8), nb), b, 0, s, out, in, in_fs, iw, ih, ws, fs);
}

// steps 2-4 of genGPUTask: outL/outA/outB = blurred channels (also the forward scratch), outPlab = blurred packed Lab;
// sb / sf : three scratch planes each (plane pitch `pp` floats)
void rd_iirblur3_run(float *outL, float *outA, float *outB, uint32_t *outPlab, const uint32_t *plab, float *sb, float *sf, size_t pp, int r, int iw, int ih,
                     int nb, size_t fs, cudaStream_t s) {
  const int n = iw * ih;
  RD_LAUNCH(kf_iir_h3, dim3(rd_cdiv(ih, 32), nb), 192, 0, s, outL, outA, outB, sb, sb + pp, sb + 2 * pp, plab, r, iw, ih, fs);
  RD_LAUNCH(kf_iir_mid3, dim3(rd_cdiv(rd_cdiv(n, 4), 256), nb), 256, 0, s, outL, outA, outB, sb, sb + pp, sb + 2 * pp, plab, r, n, fs);
  RD_LAUNCH(kf_iir_v3, dim3(rd_cdiv(iw, 128), 6, nb), 128, 0, s, sf, sf + pp, sf + 2 * pp, sb, sb + pp, sb + 2 * pp, outL, outA, outB, r, iw, ih, fs);
  RD_LAUNCH(kf_iir_fin3, dim3(rd_cdiv(rd_cdiv(n, 4), 256), nb), 256, 0, s, outL, outA, outB, outPlab, sf, sf + pp, sf + 2 * pp, sb, sb + pp, sb + 2 * pp, r, n, fs);
}

// =============================================================================================== Stage A, second half
// edgevec_f + edge_plab + thinthres_f_f_f2 (oclimgutil.cl:395-471, oclrect.c:253-258) in one kernel.  A CTA produces a
// 32x32 tile of thinned edge strength.  It stages the blurred packed-Lab tile (apron 5) and the blurred L tile (apron 2)
// in shared memory, computes the edge-magnitude tile (apron 4: the bicubic taps of the four NMS samples reach 3 pixels
// back and 4 forward) and the unit gradient of every output pixel, and takes the NMS samples from shared memory.  The
// reference's mirror border rule is applied by indexing the tiles with mirrored image coordinates.
#define ET_T 32
#define ET_PA 5                      // packed Lab apron
#define ET_MA 4                      // magnitude apron
#define ET_LA 2                      // L apron
#define ET_PW (ET_T + 2 * ET_PA)
#define ET_MW (ET_T + 2 * ET_MA)
#define ET_LW (ET_T + 2 * ET_LA)
// mirror for tile staging: positions far outside the image (never consulted) are clamped so that the load stays in bounds
__device__ __forceinline__ int mirror_safe(int x, int n) { return min(max(rd_mirror1(x, n), 0), n - 1); }
struct TileMag {                     // Plane interface of rd_bicubic: the tile already holds the mirror-extended magnitude
  const float *m; int x0, y0;
  __device__ __forceinline__ float at(int x, int y) const { return m[(y - y0) * ET_MW + (x - x0)]; }
};
// Every tile is filled for ALL its positions with the value of the mirrored image position (the reference's border rule,
// oclimgutil.cl:41-45), so the inner loops index with plain offsets and no coordinate is mirrored more than once.
//  - The packed tile is unpacked once per staged pixel into three float tiles (each pixel is a neighbour of eight
//    magnitude positions; unpacking inside rd_edge_plab_at did the conversion 8 x 3 times per position).
//  - The 5x5 derivative taps with a zero coefficient are skipped: the accumulators start at +0 and a sum is -0 only when
//    both operands are, so they are never -0 and adding the +-0 of a zero tap cannot change them.
//  - Only local maxima along the gradient need the two outer NMS samples.  They are a third of the pixels, scattered over
//    the warps, so they are queued (over the then dead Lab tiles) and finished densely, one per thread.
struct EtQueued { unsigned short idx, pad; float vx, vy, am1, ap1; };
#define ET_NPOS (ET_T * ET_T)
// TMA variant (USE_TMA): CTAs whose aproned tiles lie inside the frame fetch them with two tensor-map tile loads
// (cp.async.bulk.tensor.3d -> UTMALDG, completion on an mbarrier) issued by one thread: the packed-Lab tile as a raw 42 x 44 box that
// a second pass unpacks into the three float tiles, the L tile straight into place.  CTAs on the frame border keep the plain
// staging below: the reference mirrors coordinates there, the copy engine can only fill zeros.  Measured A/B: profiles/r04*_tma_ab.txt.
// The copy engine wants the first element of a box 16-byte aligned (measured: tools/probes/tma_probe.cu - a box starting at x = 27 raises
// "illegal instruction", x = 28 is fine), so the boxes start at the multiple of four left of the apron: packed tile bx-8 .. bx+39 (48 wide,
// the kernel uses bx-5 .. bx+36), L tile bx-4 .. bx+35 (40 wide, used: bx-2 .. bx+33).
#define ET_RAWW 48
#define ET_RAWX 8                    // the raw packed box starts ET_RAWX columns left of the tile
#define ET_LTW 40                    // pitch of the L tile when the copy engine fills it
#define ET_LTX 4
struct EtMaps { CUtensorMap p, l; };
template <bool USE_TMA>
__global__ void __launch_bounds__(256) kf_edge_thin_t(float *thin, const float *blurL, const uint32_t *blurP, const __grid_constant__ EtMaps maps, int iw, int ih, size_t fs) {
  rd_batch_z(fs, thin, blurL, blurP);
  // phase 1-2: three unpacked Lab tiles (apron 5); phase 3-4: the queue of local maxima
  __shared__ __align__(16) unsigned char lab_or_queue[(3 * ET_PW * ET_PW * 4 > ET_NPOS * (int)sizeof(EtQueued)) ? 3 * ET_PW * ET_PW * 4 : ET_NPOS * (int)sizeof(EtQueued)];
  __shared__ float sm[ET_MW * ET_MW];
  __shared__ __align__(128) float sl[USE_TMA ? ET_LW * ET_LTW : ET_LW * ET_LW];
  __shared__ __align__(128) uint32_t rawP[USE_TMA ? ET_PW * ET_RAWW : 4];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int nq;
  float (*lab)[ET_PW * ET_PW] = (float (*)[ET_PW * ET_PW])lab_or_queue;
  EtQueued *queue = (EtQueued *)lab_or_queue;
  const int bx = blockIdx.x * ET_T, by = blockIdx.y * ET_T;
  const int tx = threadIdx.x, ty = threadIdx.y;
  if (tx == 0 && ty == 0) nq = 0;
  const bool interior = USE_TMA && bx - ET_RAWX >= 0 && by - ET_PA >= 0 && bx + ET_T + ET_PA <= iw && by + ET_T + ET_PA <= ih;
  const int lpitch = interior ? ET_LTW : ET_LW, lx0 = interior ? bx - ET_LTX : bx - ET_LA;      // layout of the L tile
  if (interior) {
    if (tx == 0 && ty == 0) rd_mbar_init(&bar, 1);
    __syncthreads();
    if (tx == 0 && ty == 0) {
      rd_mbar_expect(&bar, (unsigned)(ET_PW * ET_RAWW * 4 + ET_LW * ET_LTW * 4));
      rd_tma_load3(rawP, &maps.p, bx - ET_RAWX, by - ET_PA, (int)blockIdx.z, &bar);
      rd_tma_load3(sl, &maps.l, bx - ET_LTX, by - ET_LA, (int)blockIdx.z, &bar);
    }
    rd_mbar_wait(&bar, 0);
    for (int i = ty * 32 + tx; i < ET_PW * ET_PW; i += 256) {
      const int r = i / ET_PW, c = i - r * ET_PW;
      float l, a, b;
      rd_unpacklab(rawP[r * ET_RAWW + c + (ET_RAWX - ET_PA)], l, a, b);
      lab[0][i] = l; lab[1][i] = a; lab[2][i] = b;
    }
  } else {
  // stage the packed-Lab tile (apron 5), unpacked, and the L tile (apron 2) at mirrored coordinates
  for (int r = ty; r < ET_PW; r += 8) {
    const size_t rb = (size_t)mirror_safe(by - ET_PA + r, ih) * iw;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int c = tx + 32 * h;
      if (c >= ET_PW) continue;
      float l, a, b;
      rd_unpacklab(blurP[rb + mirror_safe(bx - ET_PA + c, iw)], l, a, b);
      lab[0][r * ET_PW + c] = l; lab[1][r * ET_PW + c] = a; lab[2][r * ET_PW + c] = b;
    }
  }
  for (int r = ty; r < ET_LW; r += 8) {
    const size_t rb = (size_t)mirror_safe(by - ET_LA + r, ih) * iw;
    sl[r * ET_LW + tx] = blurL[rb + mirror_safe(bx - ET_LA + tx, iw)];
    if (tx < ET_LW - 32) sl[r * ET_LW + 32 + tx] = blurL[rb + mirror_safe(bx - ET_LA + 32 + tx, iw)];
  }
  }
  __syncthreads();
  // edge magnitude (oclimgutil.cl:422-437) on the apron-4 tile.  Position (gx, gy) of the tile stands for the image position
  // (mirror(gx), mirror(gy)); its 3x3 neighbourhood is taken around THAT position (the Lab tiles are indexed through the mirrored centre).
  for (int r = ty; r < ET_MW; r += 8) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int c = tx + h * 32;
      if (c >= ET_MW) continue;
      const int gx = bx - ET_MA + c, gy = by - ET_MA + r;
      if (gx > iw + ET_MA - 1 || gy > ih + ET_MA - 1) continue;           // beyond the apron of the last in-image pixel: never read
      const int mx = rd_mirror1(gx, iw), my = rd_mirror1(gy, ih);
      // tile coordinates of the mirrored centre: the tile spans [bx-5, bx+36], and mirroring keeps a position within 5 of where it was
      const int q = (my - (by - ET_PA)) * ET_PW + (mx - (bx - ET_PA));
      float tot = 0.0f;
#pragma unroll
      for (int ch = 0; ch < 3; ch++) {
        const float *t = lab[ch] + q;
        const float n = t[-ET_PW], w = t[-1], s = t[ET_PW], e = t[1], nw = t[-ET_PW - 1], se = t[ET_PW + 1], ne = t[-ET_PW + 1], sw = t[ET_PW - 1];
        float d = __fsub_rn(__fsub_rn(__fadd_rn(n, w), s), e);
        float acc = __fadd_rn(0.0f, __fmul_rn(__fsub_rn(nw, se), d));
        d = __fsub_rn(__fadd_rn(__fsub_rn(n, w), e), s);
        acc = __fadd_rn(acc, __fmul_rn(__fsub_rn(ne, sw), d));
        const float pos = acc > 0.0f ? acc : 0.0f;
        tot = ch == 0 ? pos : __fadd_rn(tot, pos);
      }
      sm[r * ET_MW + c] = tot > 0.0f ? __fsqrt_rn(tot) : 0.0f;
    }
  }
  __syncthreads();
  const TileMag mag = {sm, bx - ET_MA, by - ET_MA};
  constexpr float V5[25] = {-4.667f, -4.083f, 0.0f, 4.083f, 4.667f, -10.024f, -0.963f, 0.0f, 0.963f, 10.024f, -14.120f, 3.622f, 0.0f, -3.622f, 14.120f,
                            -10.024f, -0.963f, 0.0f, 0.963f, 10.024f, -4.667f, -4.083f, 0.0f, 4.083f, 4.667f};                    // RD_V5C
  const unsigned lane = tx, ltmask = (1u << lane) - 1u;
#pragma unroll 1
  for (int k = 0; k < 4; k++) {
    const int x = bx + tx, y = by + ty + k * 8;
    bool want = false;
    float2 v = make_float2(0.0f, 0.0f);
    float am1 = 0.0f, ap1 = 0.0f;
    if (x < iw && y < ih) {
      float vx = 0, vy = 0;
      const float *lc = sl + (y - (by - ET_LA)) * lpitch + (x - lx0);
#pragma unroll
      for (int yy = -2; yy <= 2; yy++)
#pragma unroll
        for (int xx = -2; xx <= 2; xx++) {
          const float s = lc[yy * lpitch + xx];
          if (V5[(xx + 2) + (yy + 2) * 5] != 0.0f) vx = __fadd_rn(vx, __fmul_rn(V5[(xx + 2) + (yy + 2) * 5], s));
          if (V5[(yy + 2) + (xx + 2) * 5] != 0.0f) vy = __fadd_rn(vy, __fmul_rn(V5[(yy + 2) + (xx + 2) * 5], s));
        }
      v = rd_edgevec_normalise(vx, vy);
      // thinthres: the outer samples only matter where the pixel is a local maximum along the gradient
      const float fx = (float)x, fy = (float)y;
      const float a0 = mag.at(x, y);
      am1 = rd_bicubic(mag, __fsub_rn(fx, v.x), __fsub_rn(fy, v.y));
      ap1 = rd_bicubic(mag, __fadd_rn(fx, v.x), __fadd_rn(fy, v.y));
      want = am1 <= a0 && a0 >= ap1;
      if (!want) thin[(size_t)y * iw + x] = 0.0f;
    }
    const unsigned b = __ballot_sync(0xffffffffu, want);
    if (b) {
      int base = 0;
      if (lane == 0) base = rd_smem_fetch_add(&nq, __popc(b));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (want) {
        EtQueued e;
        e.idx = (unsigned short)((ty + k * 8) * ET_T + tx); e.pad = 0; e.vx = v.x; e.vy = v.y; e.am1 = am1; e.ap1 = ap1;
        queue[base + __popc(b & ltmask)] = e;
      }
    }
  }
  __syncthreads();
  for (int t = ty * 32 + tx, n = nq; t < n; t += 256) {
    const EtQueued e = queue[t];
    const int x = bx + (e.idx & (ET_T - 1)), y = by + (e.idx >> 5);
    const float fx = (float)x, fy = (float)y;
    const float vx2 = __fmul_rn(2.0f, e.vx), vy2 = __fmul_rn(2.0f, e.vy);
    const float am2 = rd_bicubic(mag, __fsub_rn(fx, vx2), __fsub_rn(fy, vy2));
    const float ap2 = rd_bicubic(mag, __fadd_rn(fx, vx2), __fadd_rn(fy, vy2));
    thin[(size_t)y * iw + x] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(am2, e.am1), mag.at(x, y)), e.ap1), ap2);
  }
}
void rd_edge_thin_run(float *thin, const float *blurL, const uint32_t *blurP, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  EtMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (rd_tma_ok(blurP, iw, fs) && rd_tma_ok(blurL, iw, fs) &&
      rd_tma_make_map(&maps.p, blurP, CU_TENSOR_MAP_DATA_TYPE_UINT32, iw, ih, nb, fs, ET_RAWW, ET_PW) &&
      rd_tma_make_map(&maps.l, blurL, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, iw, ih, nb, fs, ET_LTW, ET_LW)) {
    RD_LAUNCH(kf_edge_thin_t<true>, dim3(rd_cdiv(iw, ET_T), rd_cdiv(ih, ET_T), nb), dim3(32, 8), 0, s, thin, blurL, blurP, maps, iw, ih, fs);
    return;
  }
  RD_LAUNCH(kf_edge_thin_t<false>, dim3(rd_cdiv(iw, ET_T), rd_cdiv(ih, ET_T), nb), dim3(32, 8), 0, s, thin, blurL, blurP, maps, iw, ih, fs);
}

// =============================================================================================== Stage B, string clean-up
// threshold_f_f / cast_i_f (oclrect.c:262-263) + simpleJunction + simpleConnect + stringify 0 + stringify 1
// (oclrect.cl:74-135, oclrect.c:265-272) in one kernel on byte tiles in shared memory.  Output: the cleaned 0/1 string
// image as a byte plane.  Each stage shrinks the valid region by its stencil radius: 40x40 edge tile -> 38 -> 36 -> 34 -> 32.
// Bit-plane tile (rd_bits.cuh), apron 4 = junction 1 + connect 1 + stringify 1 + 1.
#define KB1_A 4
#define KB1_R (BT_PR + 2 * KB1_A)
struct ThinPos { __device__ __forceinline__ bool operator()(float t) const { return t > 0.0f; } };
__global__ void __launch_bounds__(256) kb_strings1(uint8_t *out, const float *thin, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, thin);
  __shared__ bt_plane pa[KB1_R], pz[KB1_R], pt[KB1_R];
  const int bx0 = blockIdx.x * (32 * BT_PW), by0 = blockIdx.y * BT_PR, gy0 = by0 - KB1_A;
  // edge bitmap #1 : thin > 0 (0 outside the image)
  bt_build(pa, thin, KB1_R, bx0, gy0, iw, ih, ThinPos());
  __syncthreads();
  // simpleJunction (oclrect.cl:74): 1 + number of set neighbours, isolated pixels and the 1-px frame -> 0.
  // pz: value != 0, pt: value == 2 (exactly one neighbour)
  BT_TASKS(KB1_R) {
    BT_RC;
    uint32_t nz = 0, eq2 = 0;
    if (r >= 1 && r < KB1_R - 1) {
      const int gy = gy0 + r, gx0 = bx0 + 32 * (c - 1);
      const BtNb nb = bt_neighbours<false>(pa, r, c);
      nz = nb.centre & nb.any & (bt_rowok(gy, ih, 1) ? bt_cols(gx0, iw, 1) : 0u);
      eq2 = nz & ~nb.ge2;
    }
    pz[r][c] = nz; pt[r][c] = eq2;
  }
  __syncthreads();
  // simpleConnect (oclrect.cl:97): fill one-pixel gaps next to end pixels (value 2); 2-px frame -> 0
  BT_TASKS(KB1_R) {
    BT_RC;
    uint32_t res = 0;
    if (r >= 2 && r < KB1_R - 2) {
      const int gy = gy0 + r, gx0 = bx0 + 32 * (c - 1);
      const Bt3 tn = bt_load3(pt, r - 1, c), tm = bt_load3(pt, r, c), ts = bt_load3(pt, r + 1, c), zm = bt_load3(pz, r, c);
      const uint32_t wT = bt_w(tm, 1), eT = bt_e(tm, 1), nT = tn.c, sT = ts.c, nwT = bt_w(tn, 1), neT = bt_e(tn, 1), swT = bt_w(ts, 1), seT = bt_e(ts, 1);
      const uint32_t wZ = bt_w(zm, 1), eZ = bt_e(zm, 1), nZ = pz[r - 1][c], sZ = pz[r + 1][c];
      const uint32_t pat = (wT & eZ) | (wZ & eT) | (nT & sZ) | (nZ & sT) | (nwT & seT) | (neT & swT) | (eT & swT) | (wT & seT) | (neT & sT) | (nwT & sT);
      res = (zm.c | pat) & (bt_rowok(gy, ih, 2) ? bt_cols(gx0, iw, 2) : 0u);
    }
    pa[r][c] = res;
  }
  __syncthreads();
  // stringify mod2 = 0 then 1 (oclrect.cl:123)
  BT_TASKS(KB1_R) {
    BT_RC;
    pz[r][c] = (r >= 3 && r < KB1_R - 3) ? bt_stringify(pa, r, c, bx0 + 32 * (c - 1), gy0 + r, iw, ih, 0) : 0u;
  }
  __syncthreads();
  BT_TASKS(KB1_R) {
    BT_RC;
    pa[r][c] = (r >= 4 && r < KB1_R - 4) ? bt_stringify(pz, r, c, bx0 + 32 * (c - 1), gy0 + r, iw, ih, 1) : 0u;
  }
  __syncthreads();
  bt_store_bytes(out, pa, KB1_A, bx0, by0, iw, ih);
}
void rd_strings1_run(uint8_t *out, const float *thin, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  RD_LAUNCH(kb_strings1, dim3(rd_cdiv(iw, 32 * BT_PW), rd_cdiv(ih, BT_PR), nb), 256, 0, s, out, thin, iw, ih, fs);
}

// filterStrength(500) + threshold_i_i + cast_c_i (oclrect.c:277-284) and filterStrength(2500) + threshold_i_i
// (oclrect.c:307-312) from one read of the labels: weak mask (i8, stops the edge-preserving blur) and strong-edge bitmap
// (0/1 ints, the polyline input).  filterStrength leaves the 1-px frame alone, so there the label itself decides.
__global__ void kf_filter_masks(int8_t *weak, int *strong, const int *label, const int *str, int iw, int ih, size_t fs) {
  rd_batch_z(fs, weak, strong, label, str);
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  const int p = y * iw + x, l = label[p];
  int wk = l > 0, sg = l > 0;
  if (l > 0 && x > 0 && y > 0 && x < iw - 1 && y < ih - 1) {
    const int v = str[l];
    wk = v >= 500;
    sg = v >= 500 && v >= 2500;
  }
  weak[p] = (int8_t)wk;
  strong[p] = sg;
}
void rd_filter_masks_run(int8_t *weak, int *strong, const int *label, const int *str, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  const dim3 b(32, 8);
  RD_LAUNCH(kf_filter_masks, dim3(rd_cdiv(iw, 32), rd_cdiv(ih, 8), nb), b, 0, s, weak, strong, label, str, iw, ih, fs);
}

// quantize(24,24,24) + despeckle (oclrect.cl:207-244, oclrect.c:300-303): the 34x34 quantised tile lives in shared memory.
// - quantize maps each channel through round(v * 24) / 24 in float and back to its integer code; with 4096 + 1024 possible
//   codes that is a table, filled once per device by kq_tables with the very arithmetic of the operator kernel.
// - despeckle picks, for an edge pixel, the 3x3 non-edge neighbour at the smallest distance() in Lab.  Unpacked channels
//   are (code + 0.5) / 4096 or / 1024, so the differences are exact and the squared distance is N / 2^24 with the integer
//   N = dl^2 + 16 da^2 + 16 db^2 (dl, da, db: code differences).  The canonical distance (float)sqrt((double)N / 2^24) is
//   monotone in N, and below N = 2^20 two different N never round to the same float (the square roots are more than one
//   float ulp apart), so the reference's "first neighbour with a strictly smaller distance" is the first neighbour with a
//   strictly smaller N.  Pixels that see a larger N take the floating-point path.
__device__ uint16_t g_quantL[4096], g_quantA[1024];
__global__ void kq_tables() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4096) {
    const float l = __fadd_rn(__fmul_rn((float)i, 1.0f / 4096), 0.5f / 4096);
    g_quantL[i] = (uint16_t)rd_f2u_floor_sat(__fmul_rn(__fdiv_rn(roundf(__fmul_rn(l, 24.0f)), 24.0f), 4096.0f), 4095u);
  }
  if (i < 1024) {
    const float a = __fadd_rn(__fmul_rn((float)i, 1.0f / 1024), 0.5f / 1024);
    g_quantA[i] = (uint16_t)rd_f2u_floor_sat(__fmul_rn(__fdiv_rn(roundf(__fmul_rn(a, 24.0f)), 24.0f), 1024.0f), 1023u);
  }
}
#define QD_T 32
#define QD_W (QD_T + 2)
__device__ __forceinline__ uint32_t quant24(uint32_t v) {
  return ((uint32_t)g_quantA[v >> 22] << 22) | ((uint32_t)g_quantA[(v >> 12) & 1023u] << 12) | (uint32_t)g_quantL[v & 4095u];
}
// USE_TMA: interior CTAs fetch the two 34-row tiles as boxes of QD_RAWW columns starting at bx - 3 (a multiple of four: rd_tma.cuh)
#define QD_RAWW 40
#define QD_RAWX 3                    // the boxes start QD_RAWX columns left of the aproned tile (tile origin = block origin - 1)
struct QdMaps { CUtensorMap in, thin; };
template <bool USE_TMA>
__global__ void __launch_bounds__(256) kf_quant_despeckle_t(uint32_t *out, const uint32_t *in, const float *thin, const __grid_constant__ QdMaps maps, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, in, thin);
  __shared__ uint32_t q[QD_W * QD_W];
  __shared__ uint8_t e[QD_W * QD_W];                          // 1: edge pixel (thinned strength >= 1e-6), 2: outside the image
  __shared__ unsigned short queue[QD_T * QD_T];
  __shared__ __align__(128) uint32_t rawI[USE_TMA ? QD_W * QD_RAWW : 4];
  __shared__ __align__(128) float rawT[USE_TMA ? QD_W * QD_RAWW : 4];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int nq;
  const int bx = blockIdx.x * QD_T - 1, by = blockIdx.y * QD_T - 1;
  const int lane = threadIdx.x, wy = threadIdx.y;
  if (lane == 0 && wy == 0) nq = 0;
  const bool interior = USE_TMA && bx - QD_RAWX >= 0 && by >= 0 && bx + QD_W <= iw && by + QD_W <= ih;
  if (interior) {
    if (lane == 0 && wy == 0) rd_mbar_init(&bar, 1);
    __syncthreads();
    if (lane == 0 && wy == 0) {
      rd_mbar_expect(&bar, (unsigned)(2 * QD_W * QD_RAWW * 4));
      rd_tma_load3(rawI, &maps.in, bx - QD_RAWX, by, (int)blockIdx.z, &bar);
      rd_tma_load3(rawT, &maps.thin, bx - QD_RAWX, by, (int)blockIdx.z, &bar);
    }
    rd_mbar_wait(&bar, 0);
    for (int i = wy * 32 + lane; i < QD_W * QD_W; i += 256) {
      const int ty = i / QD_W, tx = i - ty * QD_W, r = ty * QD_RAWW + tx + QD_RAWX;
      q[i] = quant24(rawI[r]);
      e[i] = rawT[r] >= 1e-6f ? 1 : 0;
    }
  } else {
  // warp -> tile rows wy, wy + 8, ...; lane -> column lane, lanes 0 / 1 also columns 32 / 33
  for (int ty = wy; ty < QD_W; ty += 8) {
    const int gy = by + ty;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int tx = lane + 32 * h;
      if (tx >= QD_W) continue;
      const int gx = bx + tx, i = ty * QD_W + tx;
      if (gx >= 0 && gx < iw && gy >= 0 && gy < ih) {
        const size_t p = (size_t)gy * iw + gx;
        q[i] = quant24(in[p]);
        e[i] = thin[p] >= 1e-6f ? 1 : 0;
      } else { q[i] = 0; e[i] = 2; }
    }
  }
  }
  __syncthreads();
  // non-edge pixels keep their quantised colour; the edge pixels (a third of the frame, scattered over every warp) are
  // queued and searched densely, one per thread
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int tx = 1 + lane, ty = 1 + wy + k * 8;
    const int gx = bx + tx, gy = by + ty;
    const int i = ty * QD_W + tx;
    const bool inside = gx < iw && gy < ih;
    const bool edge = inside && e[i] == 1;
    if (inside && !edge) out[(size_t)gy * iw + gx] = q[i];
    const unsigned b = __ballot_sync(0xffffffffu, edge);
    if (b) {
      int base = 0;
      if (lane == 0) base = rd_smem_fetch_add(&nq, __popc(b));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (edge) queue[base + __popc(b & ((1u << lane) - 1u))] = (unsigned short)i;
    }
  }
  __syncthreads();
  for (int t = wy * 32 + lane, n = nq; t < n; t += 256) {
    const int i = queue[t];
    const int gx = bx + i % QD_W, gy = by + i / QD_W;
    const uint32_t c = q[i];
    uint32_t r = c;
    const int l0 = c & 4095u, a0 = (c >> 12) & 1023u, b0 = c >> 22;
    unsigned best = 0xffffffffu, worst = 0;
#pragma unroll
    for (int yy = -1; yy <= 1; yy++)
#pragma unroll
      for (int xx = -1; xx <= 1; xx++) {
        const int j = i + yy * QD_W + xx;
        if (e[j] != 0) continue;                              // edge pixels and positions outside the image do not donate
        const uint32_t v = q[j];
        const int dl = (int)(v & 4095u) - l0, da = (int)((v >> 12) & 1023u) - a0, db = (int)(v >> 22) - b0;
        const unsigned N = (unsigned)(dl * dl) + 16u * (unsigned)(da * da + db * db);
        worst = max(worst, N);
        if (N < best) { best = N; r = v; }
      }
    if (worst >= (1u << 20)) {
      // far-apart colours: distances may collide after rounding, follow the reference's float comparison
      r = c;
      float dist = 1e+10f, fl0, fa0, fb0;
      rd_unpacklab(c, fl0, fa0, fb0);
      for (int yy = -1; yy <= 1; yy++)
        for (int xx = -1; xx <= 1; xx++) {
          const int j = i + yy * QD_W + xx;
          if (e[j] != 0) continue;
          float l1, a1, b1;
          const uint32_t v = q[j];
          rd_unpacklab(v, l1, a1, b1);
          const float d = rd_distance3(__fsub_rn(l1, fl0), __fsub_rn(a1, fa0), __fsub_rn(b1, fb0));
          if (d < dist) { r = v; dist = d; }
        }
    }
    out[(size_t)gy * iw + gx] = r;
  }
}
// fills the quantisation tables of the current device (once; called while an oclrect_t is created, before any schedule runs)
void rd_quant_tables_init() {
  static std::mutex mu;
  static bool ready[64] = {false};
  int dev = 0;
  RD_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> g(mu);
  if (dev < 64 && ready[dev]) return;
  RD_LAUNCH(kq_tables, 16, 256, 0, (cudaStream_t)0);
  RD_CUDA(cudaStreamSynchronize(cudaStreamLegacy));          // (not a device-wide wait: another thread's stream may be capturing)
  if (dev < 64) ready[dev] = true;
}
void rd_quant_despeckle_run(uint32_t *out, const uint32_t *in, const float *thin, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  QdMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (rd_tma_ok(in, iw, fs) && rd_tma_ok(thin, iw, fs) && rd_tma_make_map(&maps.in, in, CU_TENSOR_MAP_DATA_TYPE_UINT32, iw, ih, nb, fs, QD_RAWW, QD_W) &&
      rd_tma_make_map(&maps.thin, thin, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, iw, ih, nb, fs, QD_RAWW, QD_W)) {
    RD_LAUNCH(kf_quant_despeckle_t<true>, dim3(rd_cdiv(iw, QD_T), rd_cdiv(ih, QD_T), nb), dim3(32, 8), 0, s, out, in, thin, maps, iw, ih, fs);
    return;
  }
  RD_LAUNCH(kf_quant_despeckle_t<false>, dim3(rd_cdiv(iw, QD_T), rd_cdiv(ih, QD_T), nb), dim3(32, 8), 0, s, out, in, thin, maps, iw, ih, fs);
}

// simpleJunction of the strong edges + clear + mkMergeMask0 + mkMergeMask1 (oclrect.cl:74, 246-287, oclrect.c:314-321).
// The reference scatters constants around every junction-map pixel; here every output pixel gathers.  Also writes the
// junction values themselves into `junc`, which the region-size histogram then accumulates on top of.
// Bit-plane tile (rd_bits.cuh), apron 8 (the largest radius).  The discs and the ring are unions of horizontal spans, so
// every source row is first dilated horizontally by each half-width that occurs, and an output row is the OR of the
// matching dilated rows above and below it.
#define KBJ_A 8
#define KBJ_R (BT_PR + 2 * KBJ_A)
struct IntPos { __device__ __forceinline__ bool operator()(int v) const { return v > 0; } };
enum { JD_E7, JD_E6, JD_E5, JD_E3, JD_A4, JD_A3, JD_A2, JD_R45, JD_R35, JD_N };
__global__ void __launch_bounds__(256) kb_junction_mask(uint8_t *mask, int *junc, const int *strong, int iw, int ih, size_t fs) {
  rd_batch_z(fs, mask, junc, strong);
  __shared__ bt_plane pa[KBJ_R], pz[KBJ_R], pt[KBJ_R];
  __shared__ bt_plane cnt[4][BT_PR];
  __shared__ bt_plane dil[JD_N][KBJ_R];
  const int bx0 = blockIdx.x * (32 * BT_PW), by0 = blockIdx.y * BT_PR, gy0 = by0 - KBJ_A;
  bt_build(pa, strong, KBJ_R, bx0, gy0, iw, ih, IntPos());
  __syncthreads();
  // simpleJunction (oclrect.cl:74): pz = junction map != 0, pt = junction map == 2; the counts of the payload rows are kept
  BT_TASKS(KBJ_R) {
    BT_RC;
    uint32_t nz = 0, eq2 = 0;
    if (r >= 1 && r < KBJ_R - 1) {
      const int gy = gy0 + r, gx0 = bx0 + 32 * (c - 1);
      const BtNb nb = bt_neighbours<false>(pa, r, c);
      nz = nb.centre & nb.any & (bt_rowok(gy, ih, 1) ? bt_cols(gx0, iw, 1) : 0u);
      eq2 = nz & ~nb.ge2;
      if (r >= KBJ_A && r < KBJ_A + BT_PR) {
        uint32_t k[4];
        bt_count8(pa, r, c, k);
#pragma unroll
        for (int b = 0; b < 4; b++) cnt[b][r - KBJ_A][c] = k[b];
      }
    }
    pz[r][c] = nz; pt[r][c] = eq2;
  }
  __syncthreads();
  // horizontal dilations of every row: sym[d] = pixels at distance exactly d to the left or right
  BT_TASKS(KBJ_R) {
    BT_RC;
    const Bt3 e = bt_load3(pt, r, c), a = bt_load3(pz, r, c);
    uint32_t es[8], as[6];
    es[0] = e.c; as[0] = a.c;
#pragma unroll
    for (int d = 1; d < 8; d++) es[d] = bt_w(e, d) | bt_e(e, d);
#pragma unroll
    for (int d = 1; d < 6; d++) as[d] = bt_w(a, d) | bt_e(a, d);
    const uint32_t e3 = es[0] | es[1] | es[2] | es[3], e5 = e3 | es[4] | es[5], e6 = e5 | es[6], e7 = e6 | es[7];
    const uint32_t a2 = as[0] | as[1] | as[2], a3 = a2 | as[3], a4 = a3 | as[4], r45 = as[4] | as[5], r35 = as[3] | r45;
    dil[JD_E7][r][c] = e7; dil[JD_E6][r][c] = e6; dil[JD_E5][r][c] = e5; dil[JD_E3][r][c] = e3;
    dil[JD_A4][r][c] = a4; dil[JD_A3][r][c] = a3; dil[JD_A2][r][c] = a2; dil[JD_R45][r][c] = r45; dil[JD_R35][r][c] = r35;
  }
  __syncthreads();
  // mkMergeMask0 / mkMergeMask1 (oclrect.cl:246-287): ring 16 <= d^2 < 36 around junction-map pixels := 1, then
  // disc d^2 < 64 around end pixels (== 2) and d^2 < 16 around the others := 0.  Half-widths per |dy|:
  // disc 8: 7 7 7 7 6 6 5 3, disc 4: 3 3 3 2, ring: |dx| in [4,5] [4,5] [4,5] [3,5] [0,4] [0,3]
  {
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    const int c = 1 + (lane >> 3);
    for (int pr = wy; pr < BT_PR; pr += 8) {
      const int gy = by0 + pr, r = KBJ_A + pr;
      if (gy >= ih) break;
      uint32_t zero = 0, one = 0;
#pragma unroll
      for (int dy = -7; dy <= 7; dy++) {
        const int ad = dy < 0 ? -dy : dy;
        zero |= dil[ad <= 3 ? JD_E7 : (ad <= 5 ? JD_E6 : (ad == 6 ? JD_E5 : JD_E3))][r + dy][c];
        if (ad <= 3) zero |= dil[ad <= 2 ? JD_A3 : JD_A2][r + dy][c];
        if (ad <= 5) one |= dil[ad <= 2 ? JD_R45 : (ad == 3 ? JD_R35 : (ad == 4 ? JD_A4 : JD_A3))][r + dy][c];
      }
      const uint32_t m = one & ~zero;
      const int gx = bx0 + 4 * lane, sh = (lane & 7) * 4;
      const uint32_t nz = pz[r][c] >> sh, k0 = cnt[0][pr][c] >> sh, k1 = cnt[1][pr][c] >> sh, k2 = cnt[2][pr][c] >> sh, k3 = cnt[3][pr][c] >> sh;
      int jv[4];
#pragma unroll
      for (int k = 0; k < 4; k++)
        jv[k] = ((nz >> k) & 1u) ? 1 + (int)(((k0 >> k) & 1u) | (((k1 >> k) & 1u) << 1) | (((k2 >> k) & 1u) << 2) | (((k3 >> k) & 1u) << 3)) : 0;
      const size_t p = (size_t)gy * iw + gx;
      if ((iw & 3) == 0) {
        if (gx < iw) {
          *(uint32_t *)(mask + p) = bt_nibble_bytes(m, lane);
          *(int4 *)(junc + p) = make_int4(jv[0], jv[1], jv[2], jv[3]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (gx + k < iw) { mask[p + k] = (uint8_t)((m >> (sh + k)) & 1u); junc[p + k] = jv[k]; }
      }
    }
  }
}
void rd_junction_mask_run(uint8_t *mask, int *junc, const int *strong, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  RD_LAUNCH(kb_junction_mask, dim3(rd_cdiv(iw, 32 * BT_PW), rd_cdiv(ih, BT_PR), nb), 256, 0, s, mask, junc, strong, iw, ih, fs);
}
