// rd_imgutil.cu - Stage A operators (oclimgutil.h:74-100) as sm_100a kernels.
//
// One exported function per reference wrapper (oclimgutil.c:140-319), same argument meaning, asynchronous on the
// queue's stream.  All float arithmetic follows the canonical rules of DESIGN.md / SURVEY.md section 9: IEEE
// binary32 mul/add/div/sqrt, no FMA contraction (this file is compiled with -fmad=false), rsqrt(x) := 1/sqrt(x).
// The rect pipeline (rd_rect.cu) uses fused variants of several of these kernels; the ones here are the
// operator-level drop-ins that poly.cpp / vidpoly.cpp style callers use directly.
#include "rd_common.cuh"
#define RD_TABLE_QUAL static __device__ const
#include "rd_tables.inc"
#include "rd_stageA.cuh"

struct oclimgutil_t { uint32_t magic; int ordinal; };
#define IMGUTIL_MAGIC 0xa640d893u
static inline void chk(oclimgutil_t *t) { if (!t || t->magic != IMGUTIL_MAGIC) exitf(-1, "rectdetect_b200: bad oclimgutil_t\n"); }

// ------------------------------------------------------------------------------------------ 1-D kernels
__global__ void k_clear(int *out, int n, size_t fs) {
  rd_batch_y(fs, out);                                   // oclimgutil.cl:197
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = 0;
}
__global__ void k_clear4(int4 *out, int n4, size_t fs) {                     // 128-bit stores, 4 ints per thread
  rd_batch_y(fs, out);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) out[i] = make_int4(0, 0, 0, 0);
}
__global__ void k_copy4(int4 *out, const int4 *in, int n4, size_t fs) {
  rd_batch_y(fs, out, in);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) out[i] = in[i];
}
__global__ void k_copy(int *out, const int *in, int n, size_t fs) {
  rd_batch_y(fs, out, in);                     // oclimgutil.cl:204
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
__global__ void k_cast_i_f(int *out, const float *in, float scale, int n, size_t fs) {
  rd_batch_y(fs, out, in);  // oclimgutil.cl:211
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int)__fmul_rn(in[i], scale);
}
__global__ void k_cast_c_i(int8_t *out, const int *in, int n, size_t fs) {
  rd_batch_y(fs, out, in);              // oclimgutil.cl:218
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int8_t)in[i];
}
__global__ void k_threshold_i_i(int *out, const int *in, int vlow, int thr, int vhigh, int n, size_t fs) {
  rd_batch_y(fs, out, in);   // oclimgutil.cl:225
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] > thr ? vhigh : vlow;
}
__global__ void k_threshold_f_f(float *out, const float *in, float vlow, float thr, float vhigh, int n, size_t fs) {
  rd_batch_y(fs, out, in);  // oclimgutil.cl:232
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] > thr ? vhigh : vlow;
}
__global__ void k_rand(int *out, uint64_t seed, int n, size_t fs) {
  rd_batch_y(fs, out);                     // oclimgutil.cl:248
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = rd_rand_at(i, seed);
}

// ------------------------------------------------------------------------------------------ 2-D kernels
__global__ void k_bgr2plab(uint32_t *out, const uint8_t *in, int iw, int ih, int ws, size_t fs) {
  rd_batch_z(fs, out, in);   // oclimgutil.cl:256
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  const uint8_t *p = in + (size_t)y * ws + x * 3;
  out[y * iw + x] = rd_srgb2plab(p[0], p[1], p[2], RD_S2L, RD_CFUNC, RD_CFUNC2);
}
__global__ void k_unpack_plab(float *o0, float *o1, float *o2, const uint32_t *in, int n, size_t fs) {
  rd_batch_y(fs, o0, o1, o2, in);   // oclimgutil.cl:333
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float l, a, b;
  rd_unpacklab(in[i], l, a, b);
  o0[i] = l; o1[i] = a; o2[i] = b;
}
__global__ void k_pack_plab(uint32_t *out, const float *i0, const float *i1, const float *i2, int n, size_t fs) {
  rd_batch_y(fs, out, i0, i1, i2);   // oclimgutil.cl:325
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = rd_packlab(i0[i], i1[i], i2[i]);
}

// oclimgutil.cl:395-420 ; global-memory version of the 5x5 derivative pair (the rect pipeline uses the tiled one)
__global__ void k_edgevec_f(float2 *dst, const float *in, int iw, int ih, size_t fs) {
  rd_batch_z(fs, dst, in);
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  float vx = 0, vy = 0;
  for (int yy = -2; yy <= 2; yy++)
    for (int xx = -2; xx <= 2; xx++) {
      float s = in[rd_mirror(x + xx, y + yy, iw, ih)];
      vx = __fadd_rn(vx, __fmul_rn(RD_V5C[(xx + 2) + (yy + 2) * 5], s));
      vy = __fadd_rn(vy, __fmul_rn(RD_V5C[(yy + 2) + (xx + 2) * 5], s));
    }
  dst[y * iw + x] = rd_edgevec_normalise(vx, vy);
}

// oclimgutil.cl:422-437
__global__ void k_edge_plab(float *out, const uint32_t *in, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, in);
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  out[y * iw + x] = rd_edge_plab_at(in[rd_mirror(x, y - 1, iw, ih)], in[rd_mirror(x - 1, y, iw, ih)], in[rd_mirror(x, y + 1, iw, ih)],
                                    in[rd_mirror(x + 1, y, iw, ih)], in[rd_mirror(x - 1, y - 1, iw, ih)], in[rd_mirror(x + 1, y + 1, iw, ih)],
                                    in[rd_mirror(x + 1, y - 1, iw, ih)], in[rd_mirror(x - 1, y + 1, iw, ih)]);
}

// oclimgutil.cl:456-471
struct GlobalPlane {
  const float *p; int iw, ih;
  __device__ __forceinline__ float at(int x, int y) const { return p[rd_mirror(x, y, iw, ih)]; }
};
__global__ void k_thinthres(float *out, const float *in, const float2 *vec, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, in, vec);
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  GlobalPlane pl = {in, iw, ih};
  out[y * iw + x] = rd_thinthres_at(pl, x, y, vec[y * iw + x]);
}

// ---- recursive Gaussian, oclimgutil.cl:542-637.  One thread per row / column chain (the rect pipeline uses the
// shared-memory-staged variant in rd_rect.cu).  The reference's warm-up stores for x < 0 / x >= iw land on elements
// the same chain overwrites later, so only the in-range stores are issued. ----
__global__ void k_iir_h(float *tmp0, float *tmp1, const float *ibuf, int r, int iw, int ih, size_t fs) {
  rd_batch_y(fs, tmp0, tmp1, ibuf);
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int y = t >> 1, dir = t & 1;
  if (y >= ih) return;
  const float *coef = RD_IIRCOEF[r];
  float c[15];
#pragma unroll
  for (int i = 0; i < 15; i++) c[i] = coef[i];
  rd_iir_taps tp;
  const float *row = ibuf + (size_t)y * iw;
  if (dir == 0) {
    float *o = tmp0 + (size_t)y * iw;
    for (int x = -(r + 1 + 8); x < iw; x++) { float d = tp.step(row[rd_mirror1(x, iw)], c); if (x >= 0) o[x] = d; }
  } else {
    float *o = tmp1 + (size_t)y * iw;
    for (int x = iw + (r + 1 + 8); x >= 0; x--) { float d = tp.step(row[rd_mirror1(x, iw)], c); if (x < iw) o[x] = d; }
  }
}
__global__ void k_iir_v(float *tmp0, float *tmp1, const float *obuf, int r, int iw, int ih, size_t fs) {
  rd_batch_z(fs, tmp0, tmp1, obuf);
  int x = blockIdx.x * blockDim.x + threadIdx.x, dir = blockIdx.y;
  if (x >= iw) return;
  const float *coef = RD_IIRCOEF[r];
  float c[15];
#pragma unroll
  for (int i = 0; i < 15; i++) c[i] = coef[i];
  rd_iir_taps tp;
  if (dir == 0) {
    for (int y = -(r + 1 + 8); y < ih; y++) { float d = tp.step(obuf[x + (size_t)rd_mirror1(y, ih) * iw], c); if (y >= 0) tmp0[x + (size_t)y * iw] = d; }
  } else {
    for (int y = ih + (r + 1 + 8); y >= 0; y--) { float d = tp.step(obuf[x + (size_t)rd_mirror1(y, ih) * iw], c); if (y < ih) tmp1[x + (size_t)y * iw] = d; }
  }
}
__global__ void k_iir_pass1(float *obuf, const float *tmp0, const float *tmp1, const float *ibuf, int r, int n, size_t fs) {
  rd_batch_y(fs, obuf, tmp0, tmp1, ibuf);   // oclimgutil.cl:580
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) obuf[i] = __fsub_rn(__fadd_rn(tmp1[i], tmp0[i]), __fmul_rn(ibuf[i], RD_IIRCOEF[r][0]));
}
__global__ void k_iir_pass3(float *obuf, const float *tmp0, const float *tmp1, int r, int n, size_t fs) {
  rd_batch_y(fs, obuf, tmp0, tmp1);                     // oclimgutil.cl:629
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) obuf[i] = __fsub_rn(__fadd_rn(tmp1[i], tmp0[i]), __fmul_rn(obuf[i], RD_IIRCOEF[r][0]));
}

// oclimgutil.cl:641-657 (identical to oclrect.cl:137-153) ; zero contributions are skipped (adding 0 is a no-op)
__global__ void k_calcStrength(int *out, const float *edge, const int *label, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, edge, label);
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x <= 0 || y <= 0 || x >= iw - 1 || y >= ih - 1) return;
  int p0 = y * iw + x, l = label[p0];
  if (l <= 0) return;
  float e = edge[p0];
  int v = (int)__fmul_rn(__fmul_rn(e, e), 10000.0f);
  if (v != 0) atomicAdd(out + l, v);
}
__global__ void k_filterStrength(int *labelinout, const int *str, int thre, int iw, int ih, size_t fs) {
  rd_batch_z(fs, labelinout, str);
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x <= 0 || y <= 0 || x >= iw - 1 || y >= ih - 1) return;
  int p0 = y * iw + x, l = labelinout[p0];
  if (l <= 0 || str[l] < thre) labelinout[p0] = -1;
}

// ------------------------------------------------------------------------------------------ host launchers shared with rd_rect.cu
static const int B1 = 256;
static const dim3 B2(32, 8);

void rd_k_clear(int *out, int nints, int nb, size_t fs, cudaStream_t s) {
  if (nints <= 0) return;
  const int n4 = (((uintptr_t)out & 15) == 0 && (fs & 15) == 0) ? nints / 4 : 0;
  if (n4 > 0) RD_LAUNCH(k_clear4, rd_gy(rd_cdiv(n4, B1), nb), B1, 0, s, (int4 *)out, n4, fs);
  if (nints - n4 * 4 > 0) RD_LAUNCH(k_clear, rd_gy(rd_cdiv(nints - n4 * 4, B1), nb), B1, 0, s, out + n4 * 4, nints - n4 * 4, fs);
}
void rd_k_copy(int *out, const int *in, int nints, int nb, size_t fs, cudaStream_t s) {
  if (nints <= 0) return;
  const int n4 = (((uintptr_t)out & 15) == 0 && ((uintptr_t)in & 15) == 0 && (fs & 15) == 0) ? nints / 4 : 0;
  if (n4 > 0) RD_LAUNCH(k_copy4, rd_gy(rd_cdiv(n4, B1), nb), B1, 0, s, (int4 *)out, (const int4 *)in, n4, fs);
  if (nints - n4 * 4 > 0) RD_LAUNCH(k_copy, rd_gy(rd_cdiv(nints - n4 * 4, B1), nb), B1, 0, s, out + n4 * 4, in + n4 * 4, nints - n4 * 4, fs);
}
void rd_k_rand(int *out, uint64_t seed, int n, int nb, size_t fs, cudaStream_t s) { if (n > 0) RD_LAUNCH(k_rand, rd_gy(rd_cdiv(n, B1), nb), B1, 0, s, out, seed, n, fs); }
void rd_k_iirblur(float *obuf, const float *ibuf, float *tmp0, float *tmp1, int r, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  const int n = iw * ih;
  RD_LAUNCH(k_iir_h, rd_gy(rd_cdiv(ih * 2, 64), nb), 64, 0, s, tmp0, tmp1, ibuf, r, iw, ih, fs);
  RD_LAUNCH(k_iir_pass1, rd_gy(rd_cdiv(n, B1), nb), B1, 0, s, obuf, tmp0, tmp1, ibuf, r, n, fs);
  RD_LAUNCH(k_iir_v, rd_gz(dim3(rd_cdiv(iw, 64), 2), nb), 64, 0, s, tmp0, tmp1, obuf, r, iw, ih, fs);
  RD_LAUNCH(k_iir_pass3, rd_gy(rd_cdiv(n, B1), nb), B1, 0, s, obuf, tmp0, tmp1, r, n, fs);
}

// ---- operators no configured path of the reference enqueues: the visualisers (packed Lab / float / label planes -> BGR8) and the
// alternative edge / thinning kernels.  Plain one-thread-per-pixel kernels: they are off the hot path and exist so that the
// L2 surface (oclimgutil.h:74-100) is complete. ----
__device__ __forceinline__ int rd_floor_clamp(float v, int lo, int hi) {      // clamp(convert_int_rtn(v), lo, hi); NaN -> lo
  if (!(v >= (float)lo)) return lo;
  if (v >= (float)hi) return hi;
  return (int)floorf(v);
}
__device__ __forceinline__ float rd_icfunc(float ft) {                        // oclimgutil.cl:136-142
  if (ft > 0.20689270648f) return __fmul_rn(__fmul_rn(ft, ft), ft);
  return __fmul_rn(__fsub_rn(ft, __fdiv_rn(16.0f, 116.0f)), __fdiv_rn(1.0f, 7.787f));
}
// plab2bgr (oclimgutil.cl:146-178, 264-273)
__global__ void k_plab2bgr(uint8_t *out, const uint32_t *in, int iw, int ih, int ws, size_t fs) {
  rd_batch_z(fs, out, in);
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  const float xn = 0.950456f, zn = 1.088754f;
  float l, a, b;
  rd_unpacklab(in[y * iw + x], l, a, b);
  l = __fmul_rn(l, 256.0f); a = __fmul_rn(a, 256.0f); b = __fmul_rn(b, 256.0f);
  float cy;
  if (l > 0.20689270648f) {
    cy = __fmul_rn(__fadd_rn(l, 16.0f), __fdiv_rn(1.0f, 116.0f));
    cy = __fmul_rn(__fmul_rn(cy, cy), cy);
  } else {
    cy = __fmul_rn(l, __fdiv_rn(1.0f, 903.3f));
  }
  const float fy = __fmul_rn((float)((int)RD_CFUNC[rd_floor_clamp(__fmul_rn(cy, 1024.0f), 0, 1023)] + 9039), __fdiv_rn(1.0f, 65536.0f));
  const float fz = __fsub_rn(fy, __fmul_rn(__fsub_rn(b, 128.0f), __fdiv_rn(1.0f, 200.0f)));
  const float fx = __fadd_rn(fy, __fmul_rn(__fsub_rn(a, 128.0f), __fdiv_rn(1.0f, 500.0f)));
  const float cx = __fmul_rn(rd_icfunc(fx), xn), cz = __fmul_rn(rd_icfunc(fz), zn);
  const float r = __fadd_rn(__fadd_rn(__fmul_rn(cx, 3.240479f), __fmul_rn(cy, -1.537150f)), __fmul_rn(cz, -0.498535f));
  const float g = __fadd_rn(__fadd_rn(__fmul_rn(cx, -0.969256f), __fmul_rn(cy, 1.875991f)), __fmul_rn(cz, 0.041556f));
  const float bl = __fadd_rn(__fadd_rn(__fmul_rn(cx, 0.055648f), __fmul_rn(cy, -0.204043f)), __fmul_rn(cz, 1.057311f));
  uint8_t *o = out + (size_t)y * ws + x * 3;
  o[2] = RD_L2S[rd_floor_clamp(__fmul_rn(r, 1024.0f), 0, 1023)];
  o[1] = RD_L2S[rd_floor_clamp(__fmul_rn(g, 1024.0f), 0, 1023)];
  o[0] = RD_L2S[rd_floor_clamp(__fmul_rn(bl, 1024.0f), 0, 1023)];
}
// oclimgutil.cl:283-289
__global__ void k_convert_bgr_lumaf(uint8_t *out, const float *in, float f, int iw, int ih, int ws, size_t fs) {
  rd_batch_z(fs, out, in);
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  uint8_t *o = out + (size_t)y * ws + x * 3;
  o[0] = o[1] = o[2] = (uint8_t)rd_floor_clamp(__fmul_rn(__fmul_rn(in[y * iw + x], f), 255.0f), 0, 255);
}
// oclimgutil.cl:291-322
__global__ void k_convert_bgr_labeli(uint8_t *out, const int *in, int bgc, int iw, int ih, int ws, size_t fs) {
  rd_batch_z(fs, out, in);
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  uint8_t *o = out + (size_t)y * ws + x * 3;
  const int c = in[y * iw + x];
  if (c == bgc) { o[0] = o[1] = o[2] = 0; return; }
  const int g = (int)((uint32_t)c * 1103515245u + 12345u);
  o[2] = (uint8_t)((((g & (7 << 0)) << 5) | 31) & 255);
  o[1] = (uint8_t)((((g & (7 << 3)) << 2) | 31) & 255);
  o[0] = (uint8_t)((((g & (7 << 6)) >> 1) | 31) & 255);
}
// oclimgutil.cl:439-452
__global__ void k_edge_f_f(float *out, const float *in, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, in);
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  const float n = in[rd_mirror(x, y - 1, iw, ih)], w = in[rd_mirror(x - 1, y, iw, ih)], s_ = in[rd_mirror(x, y + 1, iw, ih)], e = in[rd_mirror(x + 1, y, iw, ih)];
  const float nw = in[rd_mirror(x - 1, y - 1, iw, ih)], se = in[rd_mirror(x + 1, y + 1, iw, ih)], ne = in[rd_mirror(x + 1, y - 1, iw, ih)], sw = in[rd_mirror(x - 1, y + 1, iw, ih)];
  float t = __fsub_rn(__fsub_rn(__fadd_rn(n, w), s_), e);
  float sum = __fadd_rn(0.0f, __fmul_rn(__fsub_rn(nw, se), t));
  t = __fsub_rn(__fadd_rn(__fsub_rn(n, w), e), s_);
  sum = __fadd_rn(sum, __fmul_rn(__fsub_rn(ne, sw), t));
  out[y * iw + x] = __fsqrt_rn(sum > 0.0f ? sum : 0.0f);
}
// oclimgutil.cl:473-491
__global__ void k_thincubic(float *out, const float *in, const float2 *vec, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, in, vec);
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  GlobalPlane p = {in, iw, ih};
  const float2 v = vec[y * iw + x];
  const float fx = (float)x, fy = (float)y;
  const float vx2 = __fmul_rn(2.0f, v.x), vy2 = __fmul_rn(2.0f, v.y);
  const float am2 = rd_bicubic(p, __fsub_rn(fx, vx2), __fsub_rn(fy, vy2));
  const float am1 = rd_bicubic(p, __fsub_rn(fx, v.x), __fsub_rn(fy, v.y));
  const float a0 = p.at(x, y);
  const float ap1 = rd_bicubic(p, __fadd_rn(fx, v.x), __fadd_rn(fy, v.y));
  const float ap2 = rd_bicubic(p, __fadd_rn(fx, vx2), __fadd_rn(fy, vy2));
  const float C = 0.99f;
  const bool keep = __fmul_rn(am2, C) <= a0 && __fmul_rn(am1, C) <= a0 && a0 >= __fmul_rn(ap1, C) && a0 >= __fmul_rn(ap2, C);
  out[y * iw + x] = keep ? __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(am2, am1), a0), ap1), ap2) : 0.0f;
}
// oclimgutil.cl:354-393
__global__ void k_edgevec_plab(float2 *dst, const uint32_t *in, int iw, int ih, size_t fs) {
  rd_batch_z(fs, dst, in);
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= iw || y >= ih) return;
  float vx3[3] = {0, 0, 0}, vy3[3] = {0, 0, 0};
  for (int yy = -2; yy <= 2; yy++)
    for (int xx = -2; xx <= 2; xx++) {
      float s3[3];
      rd_unpacklab(in[rd_mirror(x + xx, y + yy, iw, ih)], s3[0], s3[1], s3[2]);
#pragma unroll
      for (int c = 0; c < 3; c++) {
        vx3[c] = __fadd_rn(vx3[c], __fmul_rn(RD_V5C[(xx + 2) + (yy + 2) * 5], s3[c]));
        vy3[c] = __fadd_rn(vy3[c], __fmul_rn(RD_V5C[(yy + 2) + (xx + 2) * 5], s3[c]));
      }
    }
  float iv3[3];
#pragma unroll
  for (int c = 0; c < 3; c++) iv3[c] = __fadd_rn(__fmul_rn(vx3[c], vx3[c]), __fmul_rn(vy3[c], vy3[c]));
  float ivlen, vx, vy;
  if (iv3[0] >= iv3[1] && iv3[0] >= iv3[2]) { ivlen = iv3[0]; vx = vx3[0]; vy = vy3[0]; }
  else if (iv3[1] >= iv3[2]) { ivlen = iv3[1]; vx = vx3[1]; vy = vy3[1]; }
  else { ivlen = iv3[2]; vx = vx3[2]; vy = vy3[2]; }
  if ((double)iv3[0] >= 1e-6 && __fadd_rn(__fmul_rn(vx3[0], vx), __fmul_rn(vy3[0], vy)) < 0.0f) { vx = -vx; vy = -vy; }
  if ((double)ivlen > 1e-10) {
    ivlen = __fdiv_rn(1.0f, __fsqrt_rn(ivlen));
    vx = __fmul_rn(vx, ivlen); vy = __fmul_rn(vy, ivlen);
  } else {
    vx = vy = 0.70710678118f;
  }
  dst[y * iw + x] = make_float2(vx, vy);
}

extern "C" {

oclimgutil_t *init_oclimgutil(cl_device_id device, cl_context) {             // oclimgutil.c:20-98
  if (rd_device_count() <= 0) exitf(-1, "rectdetect_b200: no CUDA device; there is no CPU fallback\n");
  oclimgutil_t *t = (oclimgutil_t *)calloc(1, sizeof(oclimgutil_t));
  t->magic = IMGUTIL_MAGIC;
  t->ordinal = device ? device->ordinal : 0;
  return t;
}
void dispose_oclimgutil(oclimgutil_t *thiz) { chk(thiz); thiz->magic = 0; free(thiz); }

#define OP_PROLOGUE cudaStream_t s = rd_stream(queue); chk(thiz); rd_wait_events(s, events); const int nb = 1; const size_t fs = 0; (void)nb; (void)fs

cl_event oclimgutil_clear(oclimgutil_t *thiz, cl_mem out, int size, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;
  size = (size + 3) / 4;                                                      // bytes -> ints (oclimgutil.c:142)
  rd_need(out, (size_t)size * 4, "clear");
  rd_k_clear(rd_ptr<int>(out), size, 1, 0, s);
  return rd_make_event(s, events);
}
cl_event oclimgutil_copy(oclimgutil_t *thiz, cl_mem out, cl_mem in, int size, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;
  size = (size + 3) / 4;
  rd_need(out, (size_t)size * 4, "copy"); rd_need(in, (size_t)size * 4, "copy");
  rd_k_copy(rd_ptr<int>(out), rd_ptr<int>(in), size, 1, 0, s);
  return rd_make_event(s, events);
}
cl_event oclimgutil_cast_i_f(oclimgutil_t *thiz, cl_mem out, cl_mem in, float scale, int size, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;
  rd_need(out, (size_t)size * 4, "cast_i_f"); rd_need(in, (size_t)size * 4, "cast_i_f");
  if (size > 0) RD_LAUNCH(k_cast_i_f, rd_gy(rd_cdiv(size, B1), nb), B1, 0, s, rd_ptr<int>(out), rd_ptr<float>(in), scale, size, fs);
  return rd_make_event(s, events);
}
cl_event oclimgutil_cast_c_i(oclimgutil_t *thiz, cl_mem out, cl_mem in, int size, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;
  rd_need(out, (size_t)size, "cast_c_i"); rd_need(in, (size_t)size * 4, "cast_c_i");
  if (size > 0) RD_LAUNCH(k_cast_c_i, rd_gy(rd_cdiv(size, B1), nb), B1, 0, s, rd_ptr<int8_t>(out), rd_ptr<int>(in), size, fs);
  return rd_make_event(s, events);
}
cl_event oclimgutil_threshold_i_i(oclimgutil_t *thiz, cl_mem out, cl_mem in, int vlow, int threshold, int vhigh, int size, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;
  rd_need(out, (size_t)size * 4, "threshold_i_i"); rd_need(in, (size_t)size * 4, "threshold_i_i");
  if (size > 0) RD_LAUNCH(k_threshold_i_i, rd_gy(rd_cdiv(size, B1), nb), B1, 0, s, rd_ptr<int>(out), rd_ptr<int>(in), vlow, threshold, vhigh, size, fs);
  return rd_make_event(s, events);
}
cl_event oclimgutil_threshold_f_f(oclimgutil_t *thiz, cl_mem out, cl_mem in, float vlow, float threshold, float vhigh, int size, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;
  rd_need(out, (size_t)size * 4, "threshold_f_f"); rd_need(in, (size_t)size * 4, "threshold_f_f");
  if (size > 0) RD_LAUNCH(k_threshold_f_f, rd_gy(rd_cdiv(size, B1), nb), B1, 0, s, rd_ptr<float>(out), rd_ptr<float>(in), vlow, threshold, vhigh, size, fs);
  return rd_make_event(s, events);
}
cl_event oclimgutil_rand(oclimgutil_t *thiz, cl_mem out, int size, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;
  size = (size + 3) / 4;                                                      // oclimgutil.c:181 ; seed 0 (the reference passes none)
  rd_need(out, (size_t)size * 4, "rand");
  rd_k_rand(rd_ptr<int>(out), 0, size, 1, 0, s);
  return rd_make_event(s, events);
}
cl_event oclimgutil_convert_plab_bgr(oclimgutil_t *thiz, cl_mem out, cl_mem in, int iw, int ih, int ws, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;                                                                // runs bgr2plab (Q9)
  rd_need(out, (size_t)iw * ih * 4, "convert_plab_bgr"); rd_need(in, (size_t)ws * (ih - 1) + (size_t)iw * 3, "convert_plab_bgr");
  RD_LAUNCH(k_bgr2plab, rd_gz(rd_grid2d(iw, ih, B2), nb), B2, 0, s, rd_ptr<uint32_t>(out), rd_ptr<uint8_t>(in), iw, ih, ws, fs);
  return rd_make_event(s, events);
}
cl_event oclimgutil_unpack_f_f_f_plab(oclimgutil_t *thiz, cl_mem out0, cl_mem out1, cl_mem out2, cl_mem in, int iw, int ih, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;
  const size_t P = (size_t)iw * ih * 4;
  rd_need(out0, P, "unpack"); rd_need(out1, P, "unpack"); rd_need(out2, P, "unpack"); rd_need(in, P, "unpack");
  RD_LAUNCH(k_unpack_plab, rd_gy(rd_cdiv(iw * ih, B1), nb), B1, 0, s, rd_ptr<float>(out0), rd_ptr<float>(out1), rd_ptr<float>(out2), rd_ptr<uint32_t>(in), iw * ih, fs);
  return rd_make_event(s, events);
}
cl_event oclimgutil_pack_plab_f_f_f(oclimgutil_t *thiz, cl_mem out, cl_mem in0, cl_mem in1, cl_mem in2, int iw, int ih, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;
  const size_t P = (size_t)iw * ih * 4;
  rd_need(out, P, "pack"); rd_need(in0, P, "pack"); rd_need(in1, P, "pack"); rd_need(in2, P, "pack");
  RD_LAUNCH(k_pack_plab, rd_gy(rd_cdiv(iw * ih, B1), nb), B1, 0, s, rd_ptr<uint32_t>(out), rd_ptr<float>(in0), rd_ptr<float>(in1), rd_ptr<float>(in2), iw * ih, fs);
  return rd_make_event(s, events);
}
cl_event oclimgutil_iirblur_f_f(oclimgutil_t *thiz, cl_mem obuf, cl_mem ibuf, cl_mem tmp0, cl_mem tmp1, int r, int iw, int ih, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;
  const size_t P = (size_t)iw * ih * 4;
  rd_need(obuf, P, "iirblur"); rd_need(ibuf, P, "iirblur"); rd_need(tmp0, P, "iirblur"); rd_need(tmp1, P, "iirblur");
  if (r < 0 || r >= 32) exitf(-1, "rectdetect_b200: iirblur radius %d out of range\n", r);
  rd_k_iirblur(rd_ptr<float>(obuf), rd_ptr<float>(ibuf), rd_ptr<float>(tmp0), rd_ptr<float>(tmp1), r, iw, ih, 1, 0, s);
  return rd_make_event(s, events);
}
cl_event oclimgutil_edgevec_f2_f(oclimgutil_t *thiz, cl_mem out, cl_mem in, int iw, int ih, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;
  rd_need(out, (size_t)iw * ih * 8, "edgevec"); rd_need(in, (size_t)iw * ih * 4, "edgevec");
  RD_LAUNCH(k_edgevec_f, rd_gz(rd_grid2d(iw, ih, B2), nb), B2, 0, s, rd_ptr<float2>(out), rd_ptr<float>(in), iw, ih, fs);
  return rd_make_event(s, events);
}
cl_event oclimgutil_edge_f_plab(oclimgutil_t *thiz, cl_mem out, cl_mem in, int iw, int ih, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;
  rd_need(out, (size_t)iw * ih * 4, "edge_f_plab"); rd_need(in, (size_t)iw * ih * 4, "edge_f_plab");
  RD_LAUNCH(k_edge_plab, rd_gz(rd_grid2d(iw, ih, B2), nb), B2, 0, s, rd_ptr<float>(out), rd_ptr<uint32_t>(in), iw, ih, fs);
  return rd_make_event(s, events);
}
cl_event oclimgutil_thinthres_f_f_f2(oclimgutil_t *thiz, cl_mem out, cl_mem in, cl_mem vec, int iw, int ih, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;
  rd_need(out, (size_t)iw * ih * 4, "thinthres"); rd_need(in, (size_t)iw * ih * 4, "thinthres"); rd_need(vec, (size_t)iw * ih * 8, "thinthres");
  RD_LAUNCH(k_thinthres, rd_gz(rd_grid2d(iw, ih, B2), nb), B2, 0, s, rd_ptr<float>(out), rd_ptr<float>(in), rd_ptr<float2>(vec), iw, ih, fs);
  return rd_make_event(s, events);
}
cl_event oclimgutil_label8x_int_int(oclimgutil_t *thiz, cl_mem out, cl_mem in, cl_mem tmp, int bgc, int iw, int ih, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;                                                                // converged labels; tmp holds the per-pixel link bytes
  rd_need(out, (size_t)iw * ih * 4, "label8x"); rd_need(in, (size_t)iw * ih * 4, "label8x"); rd_need(tmp, (size_t)iw * ih, "label8x tmp");
  rd_label8x(rd_ptr<int>(out), rd_ptr<int>(in), tmp->dptr, bgc, iw, ih, 1, 0, s);
  return rd_make_event(s, events);
}
cl_event oclimgutil_calcStrength(oclimgutil_t *thiz, cl_mem out, cl_mem edge, cl_mem label, int iw, int ih, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;
  const size_t P = (size_t)iw * ih * 4;
  rd_need(out, P, "calcStrength"); rd_need(edge, P, "calcStrength"); rd_need(label, P, "calcStrength");
  RD_LAUNCH(k_calcStrength, rd_gz(rd_grid2d(iw, ih, B2), nb), B2, 0, s, rd_ptr<int>(out), rd_ptr<float>(edge), rd_ptr<int>(label), iw, ih, fs);
  return NULL;                                                                // oclimgutil.c:311-314 always passes NULL events
}
cl_event oclimgutil_filterStrength(oclimgutil_t *thiz, cl_mem labelinout, cl_mem str, int thre, int iw, int ih, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;
  const size_t P = (size_t)iw * ih * 4;
  rd_need(labelinout, P, "filterStrength"); rd_need(str, P, "filterStrength");
  RD_LAUNCH(k_filterStrength, rd_gz(rd_grid2d(iw, ih, B2), nb), B2, 0, s, rd_ptr<int>(labelinout), rd_ptr<int>(str), thre, iw, ih, fs);
  return NULL;
}

// Operators no configured path of the reference enqueues (visualisers and alternative kernels).
cl_event oclimgutil_convert_bgr_lumaf(oclimgutil_t *thiz, cl_mem out, cl_mem in, float f, int iw, int ih, int ws, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;                                                                // oclimgutil.c:191
  rd_need(out, (size_t)ws * ih, "convert_bgr_lumaf"); rd_need(in, (size_t)iw * ih * 4, "convert_bgr_lumaf");
  RD_LAUNCH(k_convert_bgr_lumaf, rd_gz(rd_grid2d(iw, ih, B2), nb), B2, 0, s, rd_ptr<uint8_t>(out), rd_ptr<float>(in), f, iw, ih, ws, fs);
  return rd_make_event(s, events);
}
cl_event oclimgutil_convert_bgr_labeli(oclimgutil_t *thiz, cl_mem out, cl_mem in, int bgc, int iw, int ih, int ws, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;                                                                // oclimgutil.c:197
  rd_need(out, (size_t)ws * ih, "convert_bgr_labeli"); rd_need(in, (size_t)iw * ih * 4, "convert_bgr_labeli");
  RD_LAUNCH(k_convert_bgr_labeli, rd_gz(rd_grid2d(iw, ih, B2), nb), B2, 0, s, rd_ptr<uint8_t>(out), rd_ptr<int>(in), bgc, iw, ih, ws, fs);
  return rd_make_event(s, events);
}
cl_event oclimgutil_edge_f_f(oclimgutil_t *thiz, cl_mem out, cl_mem in, int iw, int ih, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;                                                                // oclimgutil.c:203
  rd_need(out, (size_t)iw * ih * 4, "edge_f_f"); rd_need(in, (size_t)iw * ih * 4, "edge_f_f");
  RD_LAUNCH(k_edge_f_f, rd_gz(rd_grid2d(iw, ih, B2), nb), B2, 0, s, rd_ptr<float>(out), rd_ptr<float>(in), iw, ih, fs);
  return rd_make_event(s, events);
}
cl_event oclimgutil_thincubic_f_f_f2(oclimgutil_t *thiz, cl_mem out, cl_mem in, cl_mem vec, int iw, int ih, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;                                                                // oclimgutil.c:221
  rd_need(out, (size_t)iw * ih * 4, "thincubic"); rd_need(in, (size_t)iw * ih * 4, "thincubic"); rd_need(vec, (size_t)iw * ih * 8, "thincubic");
  RD_LAUNCH(k_thincubic, rd_gz(rd_grid2d(iw, ih, B2), nb), B2, 0, s, rd_ptr<float>(out), rd_ptr<float>(in), rd_ptr<float2>(vec), iw, ih, fs);
  return rd_make_event(s, events);
}
cl_event oclimgutil_convert_bgr_plab(oclimgutil_t *thiz, cl_mem out, cl_mem in, int iw, int ih, int ws, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;                                                                // oclimgutil.c:281-285 runs plab2bgr (names swapped)
  rd_need(out, (size_t)ws * ih, "convert_bgr_plab"); rd_need(in, (size_t)iw * ih * 4, "convert_bgr_plab");
  RD_LAUNCH(k_plab2bgr, rd_gz(rd_grid2d(iw, ih, B2), nb), B2, 0, s, rd_ptr<uint8_t>(out), rd_ptr<uint32_t>(in), iw, ih, ws, fs);
  return rd_make_event(s, events);
}
cl_event oclimgutil_edgevec_f2_plab(oclimgutil_t *thiz, cl_mem out, cl_mem in, int iw, int ih, cl_command_queue queue, const cl_event *events) {
  OP_PROLOGUE;
  rd_need(out, (size_t)iw * ih * 8, "edgevec_f2_plab"); rd_need(in, (size_t)iw * ih * 4, "edgevec_f2_plab");
  RD_LAUNCH(k_edgevec_plab, rd_gz(rd_grid2d(iw, ih, B2), nb), B2, 0, s, rd_ptr<float2>(out), rd_ptr<uint32_t>(in), iw, ih, fs);
  return rd_make_event(s, events);
}
// oclimgutil_convert_bgr_luminancef cannot run in the reference either: its wrapper sets five kernel arguments ("MMiii", oclimgutil.c:187)
// for a kernel that takes six (out, in, float f, iw, ih, ws; oclimgutil.cl:275), so clEnqueueNDRangeKernel fails with
// CL_INVALID_KERNEL_ARGS and ce() ends the process.  Same outcome here.
cl_event oclimgutil_convert_bgr_luminancef(oclimgutil_t *, cl_mem, cl_mem, int, int, int, cl_command_queue, const cl_event *) {
  exitf(-1, "rectdetect_b200: oclimgutil_convert_bgr_luminancef: CL_INVALID_KERNEL_ARGS (the reference's wrapper passes 5 of the kernel's 6 arguments, oclimgutil.c:187)\n");
  return NULL;
}

}  // extern "C"
