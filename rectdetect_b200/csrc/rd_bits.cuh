// rd_bits.cuh - binary-image stencils on bit rows.
//
// The string clean-up chains of the path (simpleJunction, simpleConnect, stringify, removeBranch: oclrect.cl:74-135,
// oclpolyline.cl:66-147) and the merge-mask scatter (oclrect.cl:246-287) are functions of 0/1 images and of the junction
// map's two predicates "!= 0" and "== 2".  A CTA therefore keeps its tile as bit planes in shared memory - one 32-bit word
// per 32 pixels of a row - and every stage is a handful of shifts and boolean operations per WORD: one thread advances 32
// pixels at a time instead of one.
//
// Tile geometry (all three kernels): payload 128 x 32 pixels = 4 words x 32 rows; one more word on each side and A rows
// above / below hold the apron (A = total stencil radius of the chain; only the 8 pixels of an apron word next to the
// payload are loaded, which is enough for radius <= 8).  Word c of row r covers image columns bx0 + 32 (c - 1) .. + 31
// of image row by0 - A + r; bit i is column + i.
#ifndef RD_BITS_CUH
#define RD_BITS_CUH
#include "rd_common.cuh"

#define BT_WORDS 6
#define BT_PW 4
#define BT_PR 32
#define BT_APX 8
typedef uint32_t bt_plane[BT_WORDS];

// bits lo .. hi-1
__device__ __forceinline__ uint32_t bt_range(int lo, int hi) {
  lo = max(lo, 0); hi = min(hi, 32);
  if (hi <= lo) return 0u;
  return (0xffffffffu >> (32 - (hi - lo))) << lo;
}
// columns of a word starting at image column gx0 that lie at least `border` pixels inside the image
__device__ __forceinline__ uint32_t bt_cols(int gx0, int iw, int border) { return bt_range(border - gx0, iw - border - gx0); }
__device__ __forceinline__ bool bt_rowok(int gy, int ih, int border) { return gy >= border && gy < ih - border; }
struct Bt3 { uint32_t l, c, r; };
__device__ __forceinline__ Bt3 bt_load3(const bt_plane *p, int r, int c) {
  Bt3 v;
  v.c = p[r][c];
  v.l = c > 0 ? p[r][c - 1] : 0u;
  v.r = c < BT_WORDS - 1 ? p[r][c + 1] : 0u;
  return v;
}
// the plane seen from one / two pixels to the left (x - d) or right (x + d)
__device__ __forceinline__ uint32_t bt_w(const Bt3 &v, int d) { return __funnelshift_l(v.l, v.c, d); }
__device__ __forceinline__ uint32_t bt_e(const Bt3 &v, int d) { return __funnelshift_r(v.c, v.r, d); }

// build a bit plane of `rows` rows from an image plane: pred(value) per pixel, 0 outside the image.  Warp w takes rows
// w, w + 8, ...; every word is one ballot.
template <class T, class Pred>
__device__ __forceinline__ void bt_build(bt_plane *dst, const T *src, int rows, int bx0, int gy0, int iw, int ih, Pred pred) {
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  for (int r = wy; r < rows; r += 8) {
    const int gy = gy0 + r;
    const bool rowin = gy >= 0 && gy < ih;
    uint32_t mine = 0;
#pragma unroll
    for (int c = 0; c < BT_WORDS; c++) {
      const int gx = bx0 + 32 * (c - 1) + lane;
      bool v = false;
      const bool want = c == 0 ? lane >= 32 - BT_APX : (c == BT_WORDS - 1 ? lane < BT_APX : true);
      if (want && rowin && gx >= 0 && gx < iw) v = pred(src[(size_t)gy * iw + gx]);
      const uint32_t b = __ballot_sync(0xffffffffu, v);
      if (lane == c) mine = b;
    }
    if (lane < BT_WORDS) dst[r][lane] = mine;
  }
}

// neighbour statistics of a 0/1 plane at (r, c): any = at least one of the 8 neighbours set, ge2 / ge3 = at least two / three
struct BtNb { uint32_t centre, any, ge2, ge3; };
template <bool GE3>
__device__ __forceinline__ BtNb bt_neighbours(const bt_plane *a, int r, int c) {
  const Bt3 n = bt_load3(a, r - 1, c), m = bt_load3(a, r, c), s = bt_load3(a, r + 1, c);
  const uint32_t nb[8] = {bt_w(n, 1), n.c, bt_e(n, 1), bt_w(m, 1), bt_e(m, 1), bt_w(s, 1), s.c, bt_e(s, 1)};
  BtNb o;
  o.centre = m.c; o.any = nb[0]; o.ge2 = 0; o.ge3 = 0;
#pragma unroll
  for (int k = 1; k < 8; k++) {
    if (GE3) o.ge3 |= o.ge2 & nb[k];
    o.ge2 |= o.any & nb[k];
    o.any |= nb[k];
  }
  return o;
}
// number of set neighbours as four bit planes (weights 1, 2, 4, 8)
__device__ __forceinline__ void bt_count8(const bt_plane *a, int r, int c, uint32_t (&cnt)[4]) {
  const Bt3 n = bt_load3(a, r - 1, c), m = bt_load3(a, r, c), s = bt_load3(a, r + 1, c);
  const uint32_t x0 = bt_w(n, 1), x1 = n.c, x2 = bt_e(n, 1), x3 = bt_w(m, 1), x4 = bt_e(m, 1), x5 = bt_w(s, 1), x6 = s.c, x7 = bt_e(s, 1);
  const uint32_t s1 = x0 ^ x1 ^ x2, c1 = (x0 & x1) | (x2 & (x0 ^ x1));
  const uint32_t s2 = x3 ^ x4 ^ x5, c2 = (x3 & x4) | (x5 & (x3 ^ x4));
  const uint32_t s3 = x6 ^ x7, c3 = x6 & x7;
  const uint32_t sA = s1 ^ s2 ^ s3, cA = (s1 & s2) | (s3 & (s1 ^ s2));
  const uint32_t sB = c1 ^ c2 ^ c3, cB = (c1 & c2) | (c3 & (c1 ^ c2));
  const uint32_t sC = sB ^ cA, cC = sB & cA;
  cnt[0] = sA; cnt[1] = sC; cnt[2] = cB ^ cC; cnt[3] = cB & cC;
}

// stringify (oclrect.cl:123, oclpolyline.cl:112): on the pixels of one checkerboard colour, inside the 1-px frame,
// a pixel with a vertical AND a horizontal 4-neighbour goes
__device__ __forceinline__ uint32_t bt_stringify(const bt_plane *src, int r, int c, int gx0, int gy, int iw, int ih, int pass) {
  const Bt3 m = bt_load3(src, r, c);
  const uint32_t n = src[r - 1][c], s = src[r + 1][c], w = bt_w(m, 1), e = bt_e(m, 1);
  const uint32_t par = ((gx0 + gy + pass) & 1) ? 0xaaaaaaaau : 0x55555555u;
  const uint32_t in1 = bt_rowok(gy, ih, 1) ? bt_cols(gx0, iw, 1) : 0u;
  return m.c & ~(in1 & par & (n | s) & (w | e));
}

// four consecutive bits -> four 0/1 bytes
__device__ __forceinline__ uint32_t bt_nibble_bytes(uint32_t word, int lane) { return (((word >> ((lane & 7) * 4)) & 15u) * 0x00204081u) & 0x01010101u; }

// write the payload of a bit plane as a 0/1 byte plane
__device__ __forceinline__ void bt_store_bytes(uint8_t *out, const bt_plane *p, int A, int bx0, int by0, int iw, int ih) {
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  for (int pr = wy; pr < BT_PR; pr += 8) {
    const int gy = by0 + pr;
    if (gy >= ih) break;
    const uint32_t word = p[A + pr][1 + (lane >> 3)];
    const int gx = bx0 + 4 * lane;
    if ((iw & 3) == 0) {
      if (gx < iw) *(uint32_t *)(out + (size_t)gy * iw + gx) = bt_nibble_bytes(word, lane);
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (gx + k < iw) out[(size_t)gy * iw + gx + k] = (uint8_t)((word >> ((lane & 7) * 4 + k)) & 1u);
    }
  }
}
#define BT_TASKS(rows) for (int t_ = threadIdx.x; t_ < (rows) * BT_WORDS; t_ += 256)
#define BT_RC const int r = t_ / BT_WORDS, c = t_ - r * BT_WORDS

#endif
