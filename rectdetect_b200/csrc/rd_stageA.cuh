// rd_stageA.cuh - per-pixel device functions of Stage A (oclimgutil.cl), shared by the operator kernels in
// rd_imgutil.cu and the fused kernels of the rect pipeline in rd_rect.cu.  Compiled with -fmad=false; the
// explicit __f*_rn intrinsics document (and pin) the evaluation order the CPU oracle uses.
#ifndef RD_STAGEA_CUH
#define RD_STAGEA_CUH
#include "rd_common.cuh"

// ---- oclimgutil.cl:182-193 / oclpolyline.cl:870-889 : per-pixel hash used as tie-break noise ----
__device__ __forceinline__ uint64_t rd_rotl64(uint64_t t, int n) {
  n &= 63;                                   // OpenCL shift counts are taken modulo 64, so n == 0 leaves t unchanged
  return n == 0 ? t : ((t << n) | (t >> (64 - n)));
}
__device__ __forceinline__ uint64_t rd_xrandom(uint64_t s) {
  uint64_t t = s;
  t = rd_rotl64(t, (int)(s >> 24)); t ^= 0xf3dd0fb7820fde37ULL;
  t = rd_rotl64(t, (int)(s >> 6));  t ^= 0xe6c6ac2c59e52811ULL;
  t = rd_rotl64(t, (int)(s >> 18)); t ^= 0x2fc7871fff7c5b45ULL;
  t = rd_rotl64(t, (int)(s >> 48)); t ^= 0x47c7e1f70aa4f7c5ULL;
  t = rd_rotl64(t, (int)(s >> 0));  t ^= 0x094f02b7fb9ba895ULL;
  t = rd_rotl64(t, (int)(s >> 12)); t ^= 0x89afda817e744570ULL;
  t = rd_rotl64(t, (int)(s >> 36)); t ^= 0xc7277d052c7bf14bULL;
  return t;
}
__device__ __forceinline__ int rd_rand_at(int x, uint64_t seed) {
  return (int)rd_xrandom(((uint64_t)(int64_t)x ^ 0xb21c2cb635b48285ULL) * 0x9b923b9cec745401ULL +
                         (seed ^ 0x7bb93d75a79d2f15ULL) * 0x22cab58ada573a29ULL);
}

// ---- oclimgutil.cl:106-134 : sRGB -> packed Lab in integer fixed point.  The float-literal factors of the
// reference, e.g. (int)(0.412453f*16384+0.5f), are the integers below (evaluated in binary32). ----
template <typename T16>
__device__ __forceinline__ uint32_t rd_srgb2plab(int u0 /*B*/, int u1 /*G*/, int u2 /*R*/, const T16 *s2l, const T16 *cfunc, const T16 *cfunc2) {
  const int ir = s2l[u2], ig = s2l[u1], ib = s2l[u0];
  const int cx = (((ir * 6758 + ig * 5859 + ib * 2956 + (1 << 14)) >> 15) * 34476 + (1 << 10)) >> 11;   // 34476 = (int)(32768/0.950456f+0.5f)
  const int cy = ((ir * 3484 + ig * 11717 + ib * 1182) + (1 << 10)) >> 11;
  const int cz = (((ir * 317 + ig * 1953 + ib * 15569 + (1 << 14)) >> 15) * 30097 + (1 << 10)) >> 11;   // 30097 = (int)(32768/1.088754f+0.5f)
  const int cl = ((((int)cfunc2[cy >> 8] * (256 - (cy & 255)) + (int)cfunc2[(cy >> 8) + 1] * (cy & 255)) >> 12) + 1) >> 1;
  const int fx = (int)cfunc[cx >> 8] * (256 - (cx & 255)) + (int)cfunc[(cx >> 8) + 1] * (cx & 255);
  const int fy = (int)cfunc[cy >> 8] * (256 - (cy & 255)) + (int)cfunc[(cy >> 8) + 1] * (cy & 255);
  const int fz = (int)cfunc[cz >> 8] * (256 - (cz & 255)) + (int)cfunc[(cz >> 8) + 1] * (cz & 255);
  const int fxy = (fx - fy + (1 << 7)) >> 8;
  const int fyz = (fy - fz + (1 << 7)) >> 8;
  const int ca = (fxy * 8031 + (134744072 + (1 << 17))) >> 18;
  const int cb = (fyz * 3213 + (134744072 + (1 << 17))) >> 18;
  // the clamp acts on the value reinterpreted as unsigned (oclimgutil.cl:130-132): negatives saturate high
  uint32_t ret = min((uint32_t)cb, 1023u);
  ret = (ret << 10) | min((uint32_t)ca, 1023u);
  ret = (ret << 12) | min((uint32_t)cl, 4095u);
  return ret;
}

// ---- oclimgutil.cl:346-352 : 5x5 derivative taps (double literals narrowed to float) ----
__constant__ const float RD_V5C[25] = {
  -4.667f,  -4.083f, 0.0f, 4.083f,  4.667f,
  -10.024f, -0.963f, 0.0f, 0.963f,  10.024f,
  -14.120f, 3.622f,  0.0f, -3.622f, 14.120f,
  -10.024f, -0.963f, 0.0f, 0.963f,  10.024f,
  -4.667f,  -4.083f, 0.0f, 4.083f,  4.667f,
};

// oclimgutil.cl:409-419 ; `ivlen > 1e-10` compares in double (Q15) ; rsqrt := 1/sqrt (Q17)
__device__ __forceinline__ float2 rd_edgevec_normalise(float vx, float vy) {
  float ivlen = __fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy));
  if ((double)ivlen > 1e-10) {
    ivlen = __fdiv_rn(1.0f, __fsqrt_rn(ivlen));
    return make_float2(__fmul_rn(vx, ivlen), __fmul_rn(vy, ivlen));
  }
  return make_float2(0.70710678118f, 0.70710678118f);
}

// ---- oclimgutil.cl:422-437 : edge magnitude from the 8 neighbours of a packed-Lab pixel ----
__device__ __forceinline__ float rd_edge_plab_at(uint32_t pn, uint32_t pw, uint32_t ps, uint32_t pe, uint32_t pnw, uint32_t pse, uint32_t pne, uint32_t psw) {
  float n[3], w[3], s[3], e[3], nw[3], se[3], ne[3], sw[3];
  rd_unpacklab(pn, n[0], n[1], n[2]);   rd_unpacklab(pw, w[0], w[1], w[2]);
  rd_unpacklab(ps, s[0], s[1], s[2]);   rd_unpacklab(pe, e[0], e[1], e[2]);
  rd_unpacklab(pnw, nw[0], nw[1], nw[2]); rd_unpacklab(pse, se[0], se[1], se[2]);
  rd_unpacklab(pne, ne[0], ne[1], ne[2]); rd_unpacklab(psw, sw[0], sw[1], sw[2]);
  float sum[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    float t = __fsub_rn(__fsub_rn(__fadd_rn(n[c], w[c]), s[c]), e[c]);
    float acc = __fadd_rn(0.0f, __fmul_rn(__fsub_rn(nw[c], se[c]), t));
    t = __fsub_rn(__fadd_rn(__fsub_rn(n[c], w[c]), e[c]), s[c]);
    acc = __fadd_rn(acc, __fmul_rn(__fsub_rn(ne[c], sw[c]), t));
    sum[c] = acc > 0.0f ? acc : 0.0f;
  }
  const float tot = __fadd_rn(__fadd_rn(sum[0], sum[1]), sum[2]);
  return tot > 0.0f ? __fsqrt_rn(tot) : 0.0f;
}

// ---- oclimgutil.cl:65-94 : bicubic sample ; (int)x truncates toward zero (Q6b) ----
__device__ __forceinline__ float rd_bicubicSub(float p0, float p1, float p2, float p3, float x) {
  const float v = __fsub_rn(p1, p2);
  const float w = __fsub_rn(p3, p0);
  float u = __fadd_rn(__fmul_rn(v, 3.0f), w);
  u = __fadd_rn(__fmul_rn(u, x), __fadd_rn(__fmul_rn(-4.0f, v), __fsub_rn(__fsub_rn(p0, p1), w)));
  u = __fadd_rn(__fmul_rn(u, x), __fsub_rn(p2, p0));
  u = __fadd_rn(__fmul_rn(__fmul_rn(u, x), 0.5f), p1);
  return u;
}
template <typename Plane>
__device__ __forceinline__ float rd_bicubic(const Plane &p, float x, float y) {
  const int ix = (int)x, iy = (int)y;
  const float fx = __fsub_rn(x, (float)ix), fy = __fsub_rn(y, (float)iy);
  float r[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int yy = iy - 1 + j;
    r[j] = rd_bicubicSub(p.at(ix - 1, yy), p.at(ix, yy), p.at(ix + 1, yy), p.at(ix + 2, yy), fx);
  }
  return rd_bicubicSub(r[0], r[1], r[2], r[3], fy);
}
// oclimgutil.cl:456-471 ; Plane::at(x, y) returns the mirrored sample
template <typename Plane>
__device__ __forceinline__ float rd_thinthres_at(const Plane &p, int x, int y, float2 v) {
  const float fx = (float)x, fy = (float)y;
  const float vx2 = __fmul_rn(2.0f, v.x), vy2 = __fmul_rn(2.0f, v.y);
  const float am2 = rd_bicubic(p, __fsub_rn(fx, vx2), __fsub_rn(fy, vy2));
  const float am1 = rd_bicubic(p, __fsub_rn(fx, v.x), __fsub_rn(fy, v.y));
  const float a0 = p.at(x, y);
  const float ap1 = rd_bicubic(p, __fadd_rn(fx, v.x), __fadd_rn(fy, v.y));
  const float ap2 = rd_bicubic(p, __fadd_rn(fx, vx2), __fadd_rn(fy, vy2));
  return (am1 <= a0 && a0 >= ap1) ? __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(am2, am1), a0), ap1), ap2) : 0.0f;
}

// ---- oclimgutil.cl:549-558 : one step of the 8-tap FIR + 7-tap feedback recurrence ----
struct rd_iir_taps {
  float i1, i2, i3, i4, i5, i6, i7;      // previous inputs
  float t0, t1, t2, t3, t4, t5, t6;      // previous outputs
  __device__ __forceinline__ rd_iir_taps() : i1(0), i2(0), i3(0), i4(0), i5(0), i6(0), i7(0), t0(0), t1(0), t2(0), t3(0), t4(0), t5(0), t6(0) {}
  __device__ __forceinline__ float step(float in, const float *c) {
    float d = __fmul_rn(in, c[0]);
    float a = __fmul_rn(c[1], i1);
    a = __fadd_rn(a, __fmul_rn(c[2], i2));
    a = __fadd_rn(a, __fmul_rn(c[3], i3));
    a = __fadd_rn(a, __fmul_rn(c[4], i4));
    a = __fadd_rn(a, __fmul_rn(c[5], i5));
    a = __fadd_rn(a, __fmul_rn(c[6], i6));
    a = __fadd_rn(a, __fmul_rn(c[7], i7));
    d = __fadd_rn(d, a);
    float b = __fmul_rn(c[8], t0);
    b = __fadd_rn(b, __fmul_rn(c[9], t1));
    b = __fadd_rn(b, __fmul_rn(c[10], t2));
    b = __fadd_rn(b, __fmul_rn(c[11], t3));
    b = __fadd_rn(b, __fmul_rn(c[12], t4));
    b = __fadd_rn(b, __fmul_rn(c[13], t5));
    b = __fadd_rn(b, __fmul_rn(c[14], t6));
    d = __fadd_rn(d, b);
    i7 = i6; i6 = i5; i5 = i4; i4 = i3; i3 = i2; i2 = i1; i1 = in;
    t6 = t5; t5 = t4; t4 = t3; t3 = t2; t2 = t1; t1 = t0; t0 = d;
    return d;
  }
};

#endif
