// ora_internal.h - shared helpers of the CPU oracle (test infrastructure, see rd_oracle.h).
// Canonical float arithmetic: IEEE binary32 +,-,*,/ and sqrt, no FMA
// contraction (the build uses -ffp-contract=off), rsqrt(x) := 1.0f/sqrtf(x), hypot/distance via an
// exact double sum and a double sqrt rounded once to float, convert_uint_rtn saturating at 0.
#ifndef ORA_INTERNAL_H
#define ORA_INTERNAL_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include "rd_oracle.h"

namespace ora {

extern ora_stats_t g_stats;

static inline double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static inline int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }  // OpenCL clamp = min(max(x,lo),hi)
static inline int cl_clamp(int x, int lo, int hi) { int t = x > lo ? x : lo; return t < hi ? t : hi; }

// oclimgutil.cl:41-63 (identical copies in oclrect.cl:19)
static inline int mirror1(int x, int iw) { return cl_clamp(x, -x, iw * 2 - 2 - x); }
static inline int mirror(int x, int y, int iw, int ih) { return mirror1(x, iw) + mirror1(y, ih) * iw; }
static inline int repeat1(int x, int iw) {
  x = x < 0 ? x + iw : x;
  x = x >= iw ? x - iw : x;
  return x;
}

// convert_uint_rtn + clamp (oclimgutil.cl:30-32).  CANONICAL (Q14): negative / NaN inputs saturate to 0.
static inline uint32_t f2u_floor_sat(float v, uint32_t hi) {
  if (!(v > 0.0f)) return 0u;
  float f = floorf(v);
  if (f >= (float)hi) return hi;
  return (uint32_t)f;
}

// oclimgutil.cl:28-34
static inline uint32_t packlab(float l, float a, float b) {
  uint32_t ret = f2u_floor_sat(b * 1024, 1023u);
  ret = (ret << 10) | f2u_floor_sat(a * 1024, 1023u);
  ret = (ret << 12) | f2u_floor_sat(l * 4096, 4095u);
  return ret;
}

// oclimgutil.cl:36-39
static inline void unpacklab(uint32_t plab, float &l, float &a, float &b) {
  l = (float)(int)(plab & 4095) * (1.0f / 4096) + (0.5f / 4096);
  a = (float)(int)((plab >> 12) & 1023) * (1.0f / 1024) + (0.5f / 1024);
  b = (float)(int)((plab >> 22) & 1023) * (1.0f / 1024) + (0.5f / 1024);
}

// oclimgutil.cl:65-74
static inline float bicubicSub(float p0, float p1, float p2, float p3, float x) {
  float u, v, w;
  v = p1 - p2;
  w = p3 - p0;
  u = v * 3.0f + w;
  u = u * x + (-4.0f * v + (p0 - p1 - w));
  u = u * x + (p2 - p0);
  u = u * x * 0.5f + p1;
  return u;
}

// CANONICAL (Q17): hypot / distance = correctly rounded from a double evaluation
static inline float hypot_c(float dx, float dy) { return (float)sqrt((double)dx * dx + (double)dy * dy); }
static inline float distance3_c(float dx, float dy, float dz) {
  return (float)sqrt((double)dx * dx + (double)dy * dy + (double)dz * dz);
}

static const int RX[8] = {1, 1, 0, -1, -1, -1, 0, 1};   // oclrect.cl:12, oclpolyline.cl:63
static const int RY[8] = {0, -1, -1, -1, 0, 1, 1, 1};

// union-find with "smaller index is the root", used by the converged label operators
struct MinUF {
  int *p;
  explicit MinUF(int *parent) : p(parent) {}
  int find(int x) const {
    while (p[x] != x) x = p[x];
    return x;
  }
  int find_compress(int x) {
    int r = find(x);
    while (p[x] != r) { int n = p[x]; p[x] = r; x = n; }
    return r;
  }
  void unite(int a, int b) {
    a = find_compress(a); b = find_compress(b);
    if (a < b) p[b] = a; else if (b < a) p[a] = b;
  }
};

// kernels implemented in the other translation units
void k_clear(int32_t *out, int nints);
void k_copy(int32_t *out, const int32_t *in, int nints);
void k_rand(int32_t *out, uint64_t seed, int size);
int  label8x(int32_t *label, const int32_t *pix, int32_t *flags, int bgc, int iw, int ih);
void k_rect_calcStrength(int32_t *out, const float *edge, const int32_t *label, int iw, int ih);
void k_rect_filterStrength(int32_t *labelinout, const int32_t *str, int thre, int iw, int ih);

}  // namespace ora
#endif
