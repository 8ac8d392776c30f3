// ora_imgutil.cpp - CPU oracle, Stage A: restatement of oclimgutil.cl kernels and oclimgutil.c wrappers.
// TEST INFRASTRUCTURE ONLY (see rd_oracle.h).  Each function cites the reference lines it follows.
#include <omp.h>
#include <vector>
#include "ora_internal.h"
#include "rd_oracle_tables.inc"

namespace ora {

ora_stats_t g_stats;

// ---- oclimgutil.cl:197-237 : clear / copy / casts / thresholds (1-D kernels) ----
void k_clear(int32_t *out, int nints) {
#pragma omp parallel for schedule(static)
  for (int x = 0; x < nints; x++) out[x] = 0;
}

void k_copy(int32_t *out, const int32_t *in, int nints) {
#pragma omp parallel for schedule(static)
  for (int x = 0; x < nints; x++) out[x] = in[x];
}

// oclimgutil.cl:182-193 / oclpolyline.cl:870-881
static inline uint64_t rotl64(uint64_t t, int n) {
  n &= 63;
  return n == 0 ? t : ((t << n) | (t >> (64 - n)));   // OpenCL shifts take the count modulo 64: t<<0 | t>>64 == t | t
}

static uint64_t xrandom(uint64_t s) {
  int n;
  uint64_t t = s;
  // NOTE: in OpenCL C `t >> (64 - n)` with n == 0 shifts by 64 & 63 == 0, so the result is t | t == t.
  n = (s >> 24) & 63; t = rotl64(t, n); t ^= 0xf3dd0fb7820fde37ULL;
  n = (s >>  6) & 63; t = rotl64(t, n); t ^= 0xe6c6ac2c59e52811ULL;
  n = (s >> 18) & 63; t = rotl64(t, n); t ^= 0x2fc7871fff7c5b45ULL;
  n = (s >> 48) & 63; t = rotl64(t, n); t ^= 0x47c7e1f70aa4f7c5ULL;
  n = (s >>  0) & 63; t = rotl64(t, n); t ^= 0x094f02b7fb9ba895ULL;
  n = (s >> 12) & 63; t = rotl64(t, n); t ^= 0x89afda817e744570ULL;
  n = (s >> 36) & 63; t = rotl64(t, n); t ^= 0xc7277d052c7bf14bULL;
  return t;
}

// oclpolyline.cl:883-889 (x is a 32-bit int promoted to ulong by sign extension; x >= 0 here)
static inline int32_t rand_at(int x, uint64_t seed) {
  return (int32_t)xrandom(((uint64_t)(int64_t)x ^ 0xb21c2cb635b48285ULL) * 0x9b923b9cec745401ULL +
                          (seed ^ 0x7bb93d75a79d2f15ULL) * 0x22cab58ada573a29ULL);
}

void k_rand(int32_t *out, uint64_t seed, int size) {
#pragma omp parallel for schedule(static)
  for (int x = 0; x < size; x++) out[x] = rand_at(x, seed);
}

// ---- oclimgutil.cl:106-134 : sRGB -> packed Lab, all integer ----
static inline uint32_t srgb2plab(int u0 /*B*/, int u1 /*G*/, int u2 /*R*/) {
  const float xn = 0.950456f, zn = 1.088754f;
  int ir = RD_S2L[u2], ig = RD_S2L[u1], ib = RD_S2L[u0];

  int cx = (((ir * (int)(0.412453f * 16384 + 0.5f) + ig * (int)(0.357580f * 16384 + 0.5f) + ib * (int)(0.180423f * 16384 + 0.5f) + (1 << 14)) >> 15) * (int)(32768 / xn + 0.5f) + (1 << 10)) >> 11;
  int cy = (((ir * (int)(0.212671f * 16384 + 0.5f) + ig * (int)(0.715160f * 16384 + 0.5f) + ib * (int)(0.072169f * 16384 + 0.5f))) + (1 << 10)) >> 11;
  int cz = (((ir * (int)(0.019334f * 16384 + 0.5f) + ig * (int)(0.119193f * 16384 + 0.5f) + ib * (int)(0.950227f * 16384 + 0.5f) + (1 << 14)) >> 15) * (int)(32768 / zn + 0.5f) + (1 << 10)) >> 11;

  int cl = (((RD_CFUNC2[cy >> 8] * (256 - (cy & 255)) + RD_CFUNC2[(cy >> 8) + 1] * (cy & 255)) >> 12) + 1) >> 1;

  int fx = RD_CFUNC[cx >> 8] * (256 - (cx & 255)) + RD_CFUNC[(cx >> 8) + 1] * (cx & 255);
  int fy = RD_CFUNC[cy >> 8] * (256 - (cy & 255)) + RD_CFUNC[(cy >> 8) + 1] * (cy & 255);
  int fz = RD_CFUNC[cz >> 8] * (256 - (cz & 255)) + RD_CFUNC[(cz >> 8) + 1] * (cz & 255);

  int fxy = (fx - fy + (1 << 7)) >> 8;
  int fyz = (fy - fz + (1 << 7)) >> 8;

  int ca = ((fxy * 8031 + (134744072 + (1 << 17))) >> 18);
  int cb = ((fyz * 3213 + (134744072 + (1 << 17))) >> 18);

  // convert_uint_rtn(int) of a negative int wraps to a huge unsigned and is then clamped to the maximum;
  // the clamp is on the unsigned value (oclimgutil.cl:130-132).
  uint32_t ub = (uint32_t)cb > 1023u ? 1023u : (uint32_t)cb;
  uint32_t ua = (uint32_t)ca > 1023u ? 1023u : (uint32_t)ca;
  uint32_t ul = (uint32_t)cl > 4095u ? 4095u : (uint32_t)cl;
  uint32_t ret = ub;
  ret = (ret << 10) | ua;
  ret = (ret << 12) | ul;
  return ret;
}

// oclimgutil.cl:256-262
static void k_bgr2plab(uint32_t *out, const uint8_t *in, int iw, int ih, int ws) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x, p1 = y * ws + x * 3;
      out[p0] = srgb2plab(in[p1 + 0], in[p1 + 1], in[p1 + 2]);
    }
}

// oclimgutil.cl:325-342
static void k_pack_plab(uint32_t *out, const float *in0, const float *in1, const float *in2, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int p0 = 0; p0 < iw * ih; p0++) out[p0] = packlab(in0[p0], in1[p0], in2[p0]);
}

static void k_unpack_plab(float *o0, float *o1, float *o2, const uint32_t *in, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int p0 = 0; p0 < iw * ih; p0++) unpacklab(in[p0], o0[p0], o1[p0], o2[p0]);
}

// oclimgutil.cl:346-352 : double literals narrowed to float when the __constant float array is initialised
static const float V5C[25] = {
  (float)-4.667,  (float)-4.083, 0.0f, (float)4.083,  (float)4.667,
  (float)-10.024, (float)-0.963, 0.0f, (float)0.963,  (float)10.024,
  (float)-14.120, (float)3.622,  0.0f, (float)-3.622, (float)14.120,
  (float)-10.024, (float)-0.963, 0.0f, (float)0.963,  (float)10.024,
  (float)-4.667,  (float)-4.083, 0.0f, (float)4.083,  (float)4.667,
};

// oclimgutil.cl:395-420
static void k_edgevec_f(float *dst, const float *in, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      float vx = 0, vy = 0;
      for (int yy = -2; yy <= 2; yy++)
        for (int xx = -2; xx <= 2; xx++) {
          float s = in[mirror(x + xx, y + yy, iw, ih)];
          vx += V5C[(xx + 2) + (yy + 2) * 5] * s;
          vy += V5C[(yy + 2) + (xx + 2) * 5] * s;
        }
      float ivlen = vx * vx + vy * vy;
      if ((double)ivlen > 1e-10) {            // Q15: unsuffixed literal -> comparison in double
        ivlen = 1.0f / sqrtf(ivlen);          // CANONICAL rsqrt (Q17)
        vx *= ivlen; vy *= ivlen;
      } else {
        vx = vy = 0.70710678118f;
      }
      dst[p0 * 2 + 0] = vx;
      dst[p0 * 2 + 1] = vy;
    }
}

// oclimgutil.cl:422-437
static void k_edge_plab(float *out, const uint32_t *in, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      float n[3], w[3], s[3], e[3], nw[3], se[3], ne[3], sw[3];
      unpacklab(in[mirror(x, y - 1, iw, ih)], n[0], n[1], n[2]);
      unpacklab(in[mirror(x - 1, y, iw, ih)], w[0], w[1], w[2]);
      unpacklab(in[mirror(x, y + 1, iw, ih)], s[0], s[1], s[2]);
      unpacklab(in[mirror(x + 1, y, iw, ih)], e[0], e[1], e[2]);
      unpacklab(in[mirror(x - 1, y - 1, iw, ih)], nw[0], nw[1], nw[2]);
      unpacklab(in[mirror(x + 1, y + 1, iw, ih)], se[0], se[1], se[2]);
      unpacklab(in[mirror(x + 1, y - 1, iw, ih)], ne[0], ne[1], ne[2]);
      unpacklab(in[mirror(x - 1, y + 1, iw, ih)], sw[0], sw[1], sw[2]);
      float sum[3];
      for (int c = 0; c < 3; c++) {
        float acc = 0, t;
        t = n[c] + w[c] - s[c] - e[c];
        acc += (nw[c] - se[c]) * t;
        t = n[c] - w[c] + e[c] - s[c];
        acc += (ne[c] - sw[c]) * t;
        sum[c] = acc > 0.0f ? acc : 0.0f;     // max((float3)0, sum)
      }
      float tot = sum[0] + sum[1] + sum[2];
      out[y * iw + x] = tot > 0 ? sqrtf(tot) : 0.0f;
    }
}

// oclimgutil.cl:87-94 ; Q6b: (int)x truncates toward zero
static inline float bicubic(const float *p, float x, float y, int iw, int ih) {
  const int ix = (int)x, iy = (int)y;
  const float fx = x - ix, fy = y - iy;
  float r[4];
  for (int j = 0; j < 4; j++) {
    int yy = iy - 1 + j;
    r[j] = bicubicSub(p[mirror(ix - 1, yy, iw, ih)], p[mirror(ix, yy, iw, ih)], p[mirror(ix + 1, yy, iw, ih)], p[mirror(ix + 2, yy, iw, ih)], fx);
  }
  return bicubicSub(r[0], r[1], r[2], r[3], fy);
}

// oclimgutil.cl:456-471
static void k_thinthres(float *out, const float *in, const float *vxy, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      float vx = vxy[p0 * 2 + 0], vy = vxy[p0 * 2 + 1];
      float am2 = bicubic(in, x - 2 * vx, y - 2 * vy, iw, ih);
      float am1 = bicubic(in, x - 1 * vx, y - 1 * vy, iw, ih);
      float a0 = in[p0];
      float ap1 = bicubic(in, x + 1 * vx, y + 1 * vy, iw, ih);
      float ap2 = bicubic(in, x + 2 * vx, y + 2 * vy, iw, ih);
      out[p0] = (am1 <= a0 && a0 >= ap1) ? (am2 + am1 + a0 + ap1 + ap2) : 0.0f;
    }
}

// ---- the operators no configured path of the reference enqueues (visualisers and alternative edge kernels): restated so that the
// whole L2 surface (oclimgutil.h:74-100) has a checker; pinned to the reference's kernels in tests/test_ref_operators.py ----
static inline int floor_clamp(float v, int lo, int hi) {             // clamp(convert_int_rtn(v), lo, hi); NaN -> lo
  if (!(v >= (float)lo)) return lo;
  if (v >= (float)hi) return hi;
  return (int)floorf(v);
}
static inline float icfunc(float ft) {                                 // oclimgutil.cl:136-142
  if (ft > 0.20689270648f) return ft * ft * ft;
  return (ft - 16.0f / 116) * (1.0f / 7.787f);
}
// oclimgutil.cl:146-178 (lab2srgb) after unpacklab; u[0..2] = B, G, R
static inline void plab2srgb(uint32_t plab, uint8_t u[3]) {
  const float xn = 0.950456f, zn = 1.088754f;
  float l, a, b;
  unpacklab(plab, l, a, b);
  l *= 256; a *= 256; b *= 256;
  float cy;
  if (l > 0.20689270648f) {
    cy = (l + 16) * (1.0f / 116.0f);
    cy = cy * cy * cy;
  } else {
    cy = l * (1.0f / 903.3f);
  }
  const float fy = (RD_CFUNC[floor_clamp(cy * 1024, 0, 1023)] + 9039) * (1.0f / 65536.0f);
  const float fz = fy - (b - 128) * (1.0f / 200.0f);
  const float fx = fy + (a - 128) * (1.0f / 500.0f);
  const float cx = icfunc(fx) * xn;
  const float cz = icfunc(fz) * zn;
  const float r = cx * 3.240479f + cy * -1.537150f + cz * -0.498535f;
  const float g = cx * -0.969256f + cy * 1.875991f + cz * 0.041556f;
  const float bl = cx * 0.055648f + cy * -0.204043f + cz * 1.057311f;
  u[2] = RD_L2S[floor_clamp(r * 1024, 0, 1023)];
  u[1] = RD_L2S[floor_clamp(g * 1024, 0, 1023)];
  u[0] = RD_L2S[floor_clamp(bl * 1024, 0, 1023)];
}
// oclimgutil.cl:264-273
static void k_plab2bgr(uint8_t *out, const uint32_t *in, int iw, int ih, int ws) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) plab2srgb(in[y * iw + x], out + (size_t)y * ws + x * 3);
}
// oclimgutil.cl:283-289
static void k_convert_bgr_lumaf(uint8_t *out, const float *in, float f, int iw, int ih, int ws) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      uint8_t *o = out + (size_t)y * ws + x * 3;
      o[0] = o[1] = o[2] = (uint8_t)floor_clamp(in[y * iw + x] * f * 255, 0, 255);
    }
}
// oclimgutil.cl:291-322
static void k_convert_bgr_labeli(uint8_t *out, const int32_t *in, int bgc, int iw, int ih, int ws) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      uint8_t *o = out + (size_t)y * ws + x * 3;
      const int c = in[y * iw + x];
      if (c == bgc) { o[0] = o[1] = o[2] = 0; continue; }
      const int g = (int)((uint32_t)c * 1103515245u + 12345u);
      o[2] = (uint8_t)((((g & (7 << 0)) << 5) | 31) & 255);
      o[1] = (uint8_t)((((g & (7 << 3)) << 2) | 31) & 255);
      o[0] = (uint8_t)((((g & (7 << 6)) >> 1) | 31) & 255);
    }
}
// oclimgutil.cl:439-452
static void k_edge_f_f(float *out, const float *in, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      float sum = 0, t;
      t = in[mirror(x, y - 1, iw, ih)] + in[mirror(x - 1, y, iw, ih)] - in[mirror(x, y + 1, iw, ih)] - in[mirror(x + 1, y, iw, ih)];
      sum += (in[mirror(x - 1, y - 1, iw, ih)] - in[mirror(x + 1, y + 1, iw, ih)]) * t;
      t = in[mirror(x, y - 1, iw, ih)] - in[mirror(x - 1, y, iw, ih)] + in[mirror(x + 1, y, iw, ih)] - in[mirror(x, y + 1, iw, ih)];
      sum += (in[mirror(x + 1, y - 1, iw, ih)] - in[mirror(x - 1, y + 1, iw, ih)]) * t;
      out[y * iw + x] = sqrtf(sum > 0.0f ? sum : 0.0f);                // sqrt(max(0.0f, sum)); fmax semantics: NaN -> 0
    }
}
// oclimgutil.cl:473-491
static void k_thincubic(float *out, const float *in, const float *vxy, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      const float vx = vxy[p0 * 2 + 0], vy = vxy[p0 * 2 + 1];
      const float am2 = bicubic(in, x - 2 * vx, y - 2 * vy, iw, ih);
      const float am1 = bicubic(in, x - 1 * vx, y - 1 * vy, iw, ih);
      const float a0 = in[p0];
      const float ap1 = bicubic(in, x + 1 * vx, y + 1 * vy, iw, ih);
      const float ap2 = bicubic(in, x + 2 * vx, y + 2 * vy, iw, ih);
      const float C = 0.99f;
      out[p0] = (am2 * C <= a0 && am1 * C <= a0 && a0 >= ap1 * C && a0 >= ap2 * C) ? (am2 + am1 + a0 + ap1 + ap2) : 0;
    }
}
// oclimgutil.cl:354-393 : the gradient direction of the Lab channel with the strongest gradient, oriented like the L gradient
static void k_edgevec_plab(float *dst, const uint32_t *in, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      float vx3[3] = {0, 0, 0}, vy3[3] = {0, 0, 0};
      for (int yy = -2; yy <= 2; yy++)
        for (int xx = -2; xx <= 2; xx++) {
          float s[3];
          unpacklab(in[mirror(x + xx, y + yy, iw, ih)], s[0], s[1], s[2]);
          for (int c = 0; c < 3; c++) {
            vx3[c] += V5C[(xx + 2) + (yy + 2) * 5] * s[c];
            vy3[c] += V5C[(yy + 2) + (xx + 2) * 5] * s[c];
          }
        }
      float iv3[3];
      for (int c = 0; c < 3; c++) iv3[c] = vx3[c] * vx3[c] + vy3[c] * vy3[c];
      float ivlen, vx, vy;
      if (iv3[0] >= iv3[1] && iv3[0] >= iv3[2]) { ivlen = iv3[0]; vx = vx3[0]; vy = vy3[0]; }
      else if (iv3[1] >= iv3[2]) { ivlen = iv3[1]; vx = vx3[1]; vy = vy3[1]; }
      else { ivlen = iv3[2]; vx = vx3[2]; vy = vy3[2]; }
      if ((double)iv3[0] >= 1e-6 && (vx3[0] * vx + vy3[0] * vy < 0)) { vx = -vx; vy = -vy; }       // Q15: double comparison
      if ((double)ivlen > 1e-10) {
        ivlen = 1.0f / sqrtf(ivlen);                                   // CANONICAL rsqrt (Q17)
        vx *= ivlen; vy *= ivlen;
      } else {
        vx = vy = 0.70710678118f;
      }
      dst[p0 * 2 + 0] = vx;
      dst[p0 * 2 + 1] = vy;
    }
}

// ---- oclimgutil.cl:542-637 : recursive Gaussian.  One chain per row / column, serial inside. ----
struct Taps {
  float iv[8], tv[8];
  Taps() { for (int i = 0; i < 8; i++) iv[i] = tv[i] = 0.0f; }
  inline float step(float in, const float *coef) {
    iv[0] = in;
    float d = iv[0] * coef[0];
    d += coef[1] * iv[1] + coef[2] * iv[2] + coef[3] * iv[3] + coef[4] * iv[4] + coef[5] * iv[5] + coef[6] * iv[6] + coef[7] * iv[7];
    d += coef[8] * tv[0] + coef[9] * tv[1] + coef[10] * tv[2] + coef[11] * tv[3] + coef[12] * tv[4] + coef[13] * tv[5] + coef[14] * tv[6];
    // iv = iv.s00123456 ; tv = tv.s00123456 ; tv.s0 = d
    for (int i = 7; i >= 1; i--) { iv[i] = iv[i - 1]; tv[i] = tv[i - 1]; }
    tv[0] = d;
    return d;
  }
};

static void k_iir_pass0a(float *tmp0, const float *ibuf, int r, int iw, int ih) {
  const int N = 8;
  const float *coef = RD_IIRCOEF[r];
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++) {
    Taps t;
    for (int x = -(r + 1 + N); x < iw; x++) {
      float d = t.step(ibuf[mirror1(x, iw) + y * iw], coef);
      tmp0[repeat1(x, iw) + y * iw] = d;
    }
  }
}

static void k_iir_pass0b(float *tmp1, const float *ibuf, int r, int iw, int ih) {
  const int N = 8;
  const float *coef = RD_IIRCOEF[r];
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++) {
    Taps t;
    for (int x = iw + (r + 1 + N); x >= 0; x--) {
      float d = t.step(ibuf[mirror1(x, iw) + y * iw], coef);
      tmp1[repeat1(x, iw) + y * iw] = d;
    }
  }
}

static void k_iir_pass1(float *obuf, float *tmp0, float *tmp1, const float *ibuf, int r, int iw, int ih) {
  const float *coef = RD_IIRCOEF[r];
#pragma omp parallel for schedule(static)
  for (int p0 = 0; p0 < iw * ih; p0++) {
    obuf[p0] = tmp1[p0] + tmp0[p0] - ibuf[p0] * coef[0];
    tmp0[p0] = tmp1[p0] = 0;
  }
}

static void k_iir_pass2a(const float *obuf, float *tmp0, int r, int iw, int ih) {
  const int N = 8;
  const float *coef = RD_IIRCOEF[r];
#pragma omp parallel for schedule(static)
  for (int x = 0; x < iw; x++) {
    Taps t;
    for (int y = -(r + 1 + N); y < ih; y++) {
      float d = t.step(obuf[x + mirror1(y, ih) * iw], coef);
      tmp0[x + repeat1(y, ih) * iw] = d;
    }
  }
}

static void k_iir_pass2b(const float *obuf, float *tmp1, int r, int iw, int ih) {
  const int N = 8;
  const float *coef = RD_IIRCOEF[r];
#pragma omp parallel for schedule(static)
  for (int x = 0; x < iw; x++) {
    Taps t;
    for (int y = ih + (r + 1 + N); y >= 0; y--) {
      float d = t.step(obuf[x + mirror1(y, ih) * iw], coef);
      tmp1[x + repeat1(y, ih) * iw] = d;
    }
  }
}

static void k_iir_pass3(float *obuf, const float *tmp0, const float *tmp1, int r, int iw, int ih) {
  const float *coef = RD_IIRCOEF[r];
#pragma omp parallel for schedule(static)
  for (int p0 = 0; p0 < iw * ih; p0++) obuf[p0] = tmp1[p0] + tmp0[p0] - obuf[p0] * coef[0];
}

// ---- oclimgutil.cl:495-538 + oclimgutil.c:227-241 : 8-connected label-equivalence CCL ----
// The reference runs labelxPreprocess + MAXPASS=10 label8xMain passes gated by dirty flags.
// CANONICAL: the fixed point of those passes = every non-background pixel carries the
// smallest linear index of its 8-connected equal-value component.  Computed here by union-find;
// ora_label8x_int_int additionally replays the reference kernel sequentially to (a) prove the fixed
// point is the same and (b) report how many passes a sequential schedule needs.
int label8x(int32_t *label, const int32_t *pix, int32_t *flags, int bgc, int iw, int ih) {
  const int n = iw * ih;
  // the reference writes its dirty flags into the first MAXPASS+1 ints of tmp (oclimgutil.cl:499-501)
  if (flags) for (int i = 0; i <= 10 && i < iw; i++) flags[i] = i == 0 ? 1 : 0;
  for (int p = 0; p < n; p++) label[p] = p;
  MinUF uf(label);
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      const int v = pix[p0];
      if (v == bgc) continue;
      if (x > 0 && pix[p0 - 1] == v) uf.unite(p0, p0 - 1);
      if (y > 0) {
        if (pix[p0 - iw] == v) uf.unite(p0, p0 - iw);
        if (x > 0 && pix[p0 - iw - 1] == v) uf.unite(p0, p0 - iw - 1);
        if (x < iw - 1 && pix[p0 - iw + 1] == v) uf.unite(p0, p0 - iw + 1);
      }
    }
  for (int p = 0; p < n; p++) label[p] = pix[p] == bgc ? -1 : uf.find_compress(p);
  __atomic_fetch_add(&g_stats.label8x_calls, 1, __ATOMIC_RELAXED);
  return 0;
}

// literal sequential replay of labelxPreprocess_int_int + label8xMain_int_int, run to a clean pass
static int label8x_reference_schedule(int32_t *label, const int32_t *pix, int bgc, int iw, int ih) {
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      if (pix[p0] == bgc) { label[p0] = -1; continue; }
      if (y > 0 && pix[p0] == pix[p0 - iw]) { label[p0] = p0 - iw; continue; }
      if (x > 0 && pix[p0] == pix[p0 - 1]) { label[p0] = p0 - 1; continue; }
      label[p0] = p0;
    }
  int pass = 0;
  for (;;) {
    pass++;
    int dirty = 0;
    for (int y = 0; y < ih; y++)
      for (int x = 0; x < iw; x++) {
        const int p0 = y * iw + x;
        int g = label[p0], og = g;
        if (g == -1) continue;
        for (int yy = -1; yy <= 1; yy++)
          for (int xx = -1; xx <= 1; xx++)
            if (0 <= x + xx && x + xx < iw && 0 <= y + yy && y + yy < ih) {
              const int p1 = (y + yy) * iw + x + xx, s = label[p1];
              if (s != -1 && s < g && pix[p0] == pix[p1]) g = s;
            }
        for (int j = 0; j < 6; j++) g = label[g];
        if (g != og) {
          if (g < label[og]) label[og] = g;
          if (g < label[p0]) label[p0] = g;
          dirty = 1;
        }
      }
    if (!dirty) break;
  }
  return pass;   // number of passes executed including the final clean one
}

// ---- oclimgutil.cl:641-657 (same code as oclrect.cl:137-153) ----
void k_rect_calcStrength(int32_t *out, const float *edge, const int32_t *label, int iw, int ih) {
  // atomic_add of ints: order independent; wrap-around is defined by using unsigned arithmetic
  for (int y = 1; y < ih - 1; y++)
    for (int x = 1; x < iw - 1; x++) {
      const int p0 = y * iw + x;
      if (label[p0] <= 0) continue;
      int v = (int)(edge[p0] * edge[p0] * 10000.0f);
      out[label[p0]] = (int32_t)((uint32_t)out[label[p0]] + (uint32_t)v);
    }
}

void k_rect_filterStrength(int32_t *labelinout, const int32_t *str, int thre, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 1; y < ih - 1; y++)
    for (int x = 1; x < iw - 1; x++) {
      const int p0 = y * iw + x;
      if (labelinout[p0] <= 0 || str[labelinout[p0]] < thre) labelinout[p0] = -1;
    }
}

}  // namespace ora

using namespace ora;

extern "C" {

void ora_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int ora_get_threads(void) { return omp_get_max_threads(); }
void ora_get_stats(ora_stats_t *out) { *out = g_stats; }
void ora_reset_stats(void) { memset(&g_stats, 0, sizeof(g_stats)); }
void ora_free(void *p) { free(p); }

// oclimgutil.c:140-145 : size is in BYTES
void ora_clear(int32_t *out, int size_bytes) { k_clear(out, (size_bytes + 3) / 4); }
void ora_copy(int32_t *out, const int32_t *in, int size_bytes) { k_copy(out, in, (size_bytes + 3) / 4); }

void ora_cast_i_f(int32_t *out, const float *in, float scale, int size) {
#pragma omp parallel for schedule(static)
  for (int x = 0; x < size; x++) out[x] = (int)(in[x] * scale);
}

void ora_cast_c_i(int8_t *out, const int32_t *in, int size) {
  // out may alias the int plane it is written into only in the reference's tmp[1] <- tmp[0] use; no overlap here
#pragma omp parallel for schedule(static)
  for (int x = 0; x < size; x++) out[x] = (int8_t)in[x];
}

void ora_threshold_i_i(int32_t *out, const int32_t *in, int vlow, int threshold, int vhigh, int size) {
#pragma omp parallel for schedule(static)
  for (int x = 0; x < size; x++) out[x] = in[x] > threshold ? vhigh : vlow;
}

void ora_threshold_f_f(float *out, const float *in, float vlow, float threshold, float vhigh, int size) {
#pragma omp parallel for schedule(static)
  for (int x = 0; x < size; x++) out[x] = in[x] > threshold ? vhigh : vlow;
}

void ora_convert_plab_bgr(uint32_t *out, const uint8_t *in, int iw, int ih, int ws) { k_bgr2plab(out, in, iw, ih, ws); }
void ora_unpack_f_f_f_plab(float *o0, float *o1, float *o2, const uint32_t *in, int iw, int ih) { k_unpack_plab(o0, o1, o2, in, iw, ih); }
void ora_pack_plab_f_f_f(uint32_t *out, const float *i0, const float *i1, const float *i2, int iw, int ih) { k_pack_plab(out, i0, i1, i2, iw, ih); }

// oclimgutil.c:243-273
void ora_iirblur_f_f(float *obuf, const float *ibuf, float *tmp0, float *tmp1, int r, int iw, int ih) {
  k_clear((int32_t *)tmp0, iw * ih);
  k_clear((int32_t *)tmp1, iw * ih);
  k_iir_pass0a(tmp0, ibuf, r, iw, ih);
  k_iir_pass0b(tmp1, ibuf, r, iw, ih);
  k_iir_pass1(obuf, tmp0, tmp1, ibuf, r, iw, ih);
  k_iir_pass2a(obuf, tmp0, r, iw, ih);
  k_iir_pass2b(obuf, tmp1, r, iw, ih);
  k_iir_pass3(obuf, tmp0, tmp1, r, iw, ih);
}

void ora_edgevec_f2_f(float *out_xy, const float *in, int iw, int ih) { k_edgevec_f(out_xy, in, iw, ih); }
// NV12 (Y plane, row stride ys; interleaved UV plane behind it) -> BGR8 rows of ws bytes: the integer BT.601 limited-range conversion
// of OpenCV's cvtColor(COLOR_YUV2BGR_NV12) (modules/imgproc/src/color_yuv.simd.hpp: ITUR_BT_601_CY 1220542, CUB 2116026, CUG -409993,
// CVG -852492, CVR 1673527, shift 20) - the step cv::VideoCapture performs in front of the reference's programs (vidrect.cpp:160-166).
// Pinned to OpenCV itself by tests/golden/nv12_golden.json (tools/make_nv12_golden.py).
void ora_nv12_to_bgr(uint8_t *bgr, const uint8_t *nv12, int iw, int ih, int ws, int ys) {
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const uint8_t *uv = nv12 + (size_t)ih * ys + (size_t)(y >> 1) * ys + (x & ~1);
      const int Y = nv12[(size_t)y * ys + x], u = (int)uv[0] - 128, v = (int)uv[1] - 128;
      const int yy = (Y - 16 > 0 ? Y - 16 : 0) * 1220542 + (1 << 19);
      const int r = (yy + 1673527 * v) >> 20, g = (yy - 852492 * v - 409993 * u) >> 20, b = (yy + 2116026 * u) >> 20;
      uint8_t *o = bgr + (size_t)y * ws + 3 * x;
      o[0] = (uint8_t)(b < 0 ? 0 : b > 255 ? 255 : b); o[1] = (uint8_t)(g < 0 ? 0 : g > 255 ? 255 : g); o[2] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
    }
}
void ora_edgevec_f2_plab(float *out_xy, const uint32_t *in, int iw, int ih) { k_edgevec_plab(out_xy, in, iw, ih); }
void ora_edge_f_f(float *out, const float *in, int iw, int ih) { k_edge_f_f(out, in, iw, ih); }
void ora_thincubic_f_f_f2(float *out, const float *in, const float *vxy, int iw, int ih) { k_thincubic(out, in, vxy, iw, ih); }
void ora_convert_bgr_plab(uint8_t *out, const uint32_t *in, int iw, int ih, int ws) { k_plab2bgr(out, in, iw, ih, ws); }     /* Q9: the names are swapped */
void ora_convert_bgr_lumaf(uint8_t *out, const float *in, float f, int iw, int ih, int ws) { k_convert_bgr_lumaf(out, in, f, iw, ih, ws); }
void ora_convert_bgr_labeli(uint8_t *out, const int32_t *in, int bgc, int iw, int ih, int ws) { k_convert_bgr_labeli(out, in, bgc, iw, ih, ws); }
void ora_edge_f_plab(float *out, const uint32_t *in, int iw, int ih) { k_edge_plab(out, in, iw, ih); }
void ora_thinthres_f_f_f2(float *out, const float *in, const float *vxy, int iw, int ih) { k_thinthres(out, in, vxy, iw, ih); }

int ora_label8x_int_int(int32_t *out, const int32_t *in, int32_t *tmp, int bgc, int iw, int ih) {
  label8x(out, in, tmp, bgc, iw, ih);
  // cross-check: the reference kernel's own fixed point must be the same labelling
  std::vector<int32_t> chk((size_t)iw * ih);
  int passes = label8x_reference_schedule(chk.data(), in, bgc, iw, ih);
  if (memcmp(chk.data(), out, sizeof(int32_t) * (size_t)iw * ih) != 0) return -passes;
  if (passes > g_stats.label8x_seq_passes) g_stats.label8x_seq_passes = passes;
  return passes;
}

void ora_calcStrength(int32_t *out, const float *edge, const int32_t *label, int iw, int ih) { k_rect_calcStrength(out, edge, label, iw, ih); }
void ora_filterStrength(int32_t *labelinout, const int32_t *str, int thre, int iw, int ih) { k_rect_filterStrength(labelinout, str, thre, iw, ih); }

uint32_t ora_srgb2plab(int b, int g, int r) { return srgb2plab(b, g, r); }
uint32_t ora_packlab(float l, float a, float b) { return packlab(l, a, b); }
void ora_unpacklab(uint32_t plab, float out[3]) { unpacklab(plab, out[0], out[1], out[2]); }
int ora_mirror1(int x, int iw) { return mirror1(x, iw); }
int ora_repeat1(int x, int iw) { return repeat1(x, iw); }
uint64_t ora_xrandom(uint64_t s) { return xrandom(s); }
int32_t ora_rand_at(int x, uint64_t seed) { return rand_at(x, seed); }

}  // extern "C"
