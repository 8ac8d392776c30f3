/* cl_compat.h - the small subset of OpenCL C that the reference's three kernel files use, as plain C++.
 *
 * TEST INFRASTRUCTURE (oracle/).  oracle/cl_translate.py turns /root/reference/{oclimgutil,oclrect,oclpolyline}.cl into
 * C++ translation units at build time (outputs under oracle/_ref/ only; the only textual change is the vector-literal
 * syntax "(float3)(a, b, c)" -> "make_float3(a, b, c)") and this header supplies everything else: address-space
 * qualifiers as empty macros, vector types with component (.x .y .z / .s0 ...) access and component-wise operators,
 * get_global_id, the conversion and math built-ins, atomics.  A work-item is a function call; an NDRange is a pair of
 * loops (oracle/_ref trampolines): in raster order on one thread - ONE legal sequential schedule of the reference's
 * kernels - or with its rows spread over OpenMP threads (ref_cl_rt.cpp).
 *
 * Built-ins whose precision OpenCL leaves to the vendor follow the canonical choices of DESIGN.md (Q14, Q16-Q18): no FP
 * contraction (-ffp-contract=off), IEEE sqrt and divide, rsqrt(x) = 1 / sqrt(x), hypot / distance = (float)sqrt(exact
 * double sum), convert_uint_rtn saturating at 0. */
#ifndef RD_CL_COMPAT_H
#define RD_CL_COMPAT_H
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long ulong;

#define __kernel
#define __global
#define global
#define __constant const
#define constant const
#define __local
#define local
#ifndef NDEBUG
#define NDEBUG 1
#endif

extern thread_local int rd_cl_gid[3];   /* ref_cl_rt.cpp */
static inline int get_global_id(int d) { return rd_cl_gid[d]; }

/* ---- vector types ---------------------------------------------------------------------------------------------------- */
template <class T> struct clv2 {
  union { struct { T x, y; }; struct { T s0, s1; }; };
  clv2() {}
  clv2(T a) : x(a), y(a) {}
  clv2(T a, T b) : x(a), y(b) {}
};
template <class T> struct clv3 {
  union { struct { T x, y, z; }; struct { T s0, s1, s2; }; };
  clv3() {}
  clv3(T a) : x(a), y(a), z(a) {}
  clv3(T a, T b, T c) : x(a), y(b), z(c) {}
};
struct float8;
struct float8_s00123456 { operator float8() const; };
struct float8 {
  union { struct { float s0, s1, s2, s3, s4, s5, s6, s7; }; float8_s00123456 s00123456; };
  float8() {}
  float8(float a) : s0(a), s1(a), s2(a), s3(a), s4(a), s5(a), s6(a), s7(a) {}
};
inline float8_s00123456::operator float8() const {
  const float *p = (const float *)this;
  float8 r;
  r.s0 = p[0]; r.s1 = p[0]; r.s2 = p[1]; r.s3 = p[2]; r.s4 = p[3]; r.s5 = p[4]; r.s6 = p[5]; r.s7 = p[6];
  return r;
}
typedef clv2<float> float2;
typedef clv3<float> float3;
typedef clv2<int> int2;
typedef clv3<int> int3;
typedef clv2<short> short2;
typedef clv3<uchar> uchar3;

#define CLV_BINOP(op)                                                                                                          \
  template <class T> inline clv2<T> operator op(clv2<T> a, clv2<T> b) { return clv2<T>(a.x op b.x, a.y op b.y); }              \
  template <class T> inline clv2<T> operator op(clv2<T> a, T b) { return clv2<T>(a.x op b, a.y op b); }                        \
  template <class T> inline clv2<T> operator op(T a, clv2<T> b) { return clv2<T>(a op b.x, a op b.y); }                        \
  template <class T> inline clv3<T> operator op(clv3<T> a, clv3<T> b) { return clv3<T>(a.x op b.x, a.y op b.y, a.z op b.z); }  \
  template <class T> inline clv3<T> operator op(clv3<T> a, T b) { return clv3<T>(a.x op b, a.y op b, a.z op b); }              \
  template <class T> inline clv3<T> operator op(T a, clv3<T> b) { return clv3<T>(a op b.x, a op b.y, a op b.z); }              \
  template <class T> inline clv2<T> &operator op##=(clv2<T> &a, clv2<T> b) { a = a op b; return a; }                           \
  template <class T> inline clv2<T> &operator op##=(clv2<T> &a, T b) { a = a op b; return a; }                                 \
  template <class T> inline clv3<T> &operator op##=(clv3<T> &a, clv3<T> b) { a = a op b; return a; }                           \
  template <class T> inline clv3<T> &operator op##=(clv3<T> &a, T b) { a = a op b; return a; }
CLV_BINOP(+)
CLV_BINOP(-)
CLV_BINOP(*)
CLV_BINOP(/)
CLV_BINOP(&)
#undef CLV_BINOP
/* mixed scalar types as OpenCL's usual arithmetic conversions would have them: int literal with a float / int vector */
inline float2 operator*(float2 a, int b) { return a * (float)b; }
inline float2 operator*(float2 a, double b) { return a * (float)b; }
inline float3 operator*(float3 a, int b) { return a * (float)b; }
inline float3 operator*(float3 a, double b) { return a * (float)b; }
inline float3 operator*(double a, float3 b) { return (float)a * b; }
inline float3 &operator*=(float3 &a, int b) { a = a * (float)b; return a; }
template <class T> inline clv2<T> operator-(clv2<T> a) { return clv2<T>(-a.x, -a.y); }
template <class T> inline clv3<T> operator-(clv3<T> a) { return clv3<T>(-a.x, -a.y, -a.z); }

#define make_float2 float2
#define make_float3 float3
#define make_int2 int2
#define make_int3 int3
#define make_short2 short2
#define make_uchar3 uchar3

/* ---- conversions ------------------------------------------------------------------------------------------------------- */
static inline int convert_int_rtn(float v) { return (int)floorf(v); }
static inline uint convert_uint_rtn(float v) { return v > 0.0f ? (v >= 4294967296.0f ? 0xffffffffu : (uint)floorf(v)) : 0u; }   /* Q14 */
static inline long convert_long_rte(float v) { return llrintf(v); }                                                              /* Q18 */
static inline int2 convert_int2(short2 v) { return int2((int)v.x, (int)v.y); }
static inline int2 convert_int2(float2 v) { return int2((int)v.x, (int)v.y); }
static inline int2 convert_int2_rte(float2 v) { return int2((int)rintf(v.x), (int)rintf(v.y)); }
static inline short2 convert_short2(float2 v) { return short2((short)(int)v.x, (short)(int)v.y); }
static inline float2 convert_float2(short2 v) { return float2((float)v.x, (float)v.y); }
static inline float2 convert_float2(int2 v) { return float2((float)v.x, (float)v.y); }
static inline float3 convert_float3(int3 v) { return float3((float)v.x, (float)v.y, (float)v.z); }
static inline float3 convert_float3(uchar3 v) { return float3((float)v.x, (float)v.y, (float)v.z); }

/* ---- math ---------------------------------------------------------------------------------------------------------------- */
static inline int clamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline uint clamp(uint v, uint lo, uint hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline float clamp(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
static inline int2 clamp(int2 v, int2 lo, int2 hi) { return int2(clamp(v.x, lo.x, hi.x), clamp(v.y, lo.y, hi.y)); }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float3 max(float3 a, float3 b) { return float3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
static inline float3 max(float3 a, float b) { return float3(fmaxf(a.x, b), fmaxf(a.y, b), fmaxf(a.z, b)); }
static inline uint abs_diff(int a, int b) { return a > b ? (uint)a - (uint)b : (uint)b - (uint)a; }
static inline float rsqrt(float x) { return 1.0f / sqrtf(x); }                                                                   /* Q17 */
static inline float cl_hypot(float a, float b) { return (float)sqrt((double)a * a + (double)b * b); }
static inline float distance(float3 a, float3 b) {
  const double dx = (double)(a.x - b.x), dy = (double)(a.y - b.y), dz = (double)(a.z - b.z);
  return (float)sqrt(dx * dx + dy * dy + dz * dz);
}
static inline float distance(float2 a, float2 b) { return cl_hypot(a.x - b.x, a.y - b.y); }
/* float overloads of the C names the kernels call with float arguments */
static inline float cl_round(float x) { return roundf(x); }
static inline float cl_sqrt(float x) { return sqrtf(x); }
static inline float cl_fabs(float x) { return fabsf(x); }
#define hypot cl_hypot
#define round cl_round
#define sqrt(x) cl_sqrt_dispatch(x)
static inline float cl_sqrt_dispatch(float x) { return sqrtf(x); }
static inline double cl_sqrt_dispatch(double x) { return __builtin_sqrt(x); }
#define fabs cl_fabs

/* ---- atomics: real ones (the NDRange may be spread over host threads, ref_cl_rt.cpp) --------------------------------------- */
template <class T> static inline T atomic_add(volatile T *p, T v) { return __atomic_fetch_add((T *)p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomic_inc(volatile T *p) { return __atomic_fetch_add((T *)p, (T)1, __ATOMIC_RELAXED); }
template <class T> static inline T atomic_min(volatile T *p, T v) {
  T o = __atomic_load_n((T *)p, __ATOMIC_RELAXED);
  while (v < o && !__atomic_compare_exchange_n((T *)p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return o;
}
template <class T> static inline T atomic_max(volatile T *p, T v) {
  T o = __atomic_load_n((T *)p, __ATOMIC_RELAXED);
  while (v > o && !__atomic_compare_exchange_n((T *)p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return o;
}
template <class T> static inline T atomic_cmpxchg(volatile T *p, T cmp, T v) {
  T o = cmp;
  __atomic_compare_exchange_n((T *)p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
  return o;
}
static inline long atom_add(volatile long *p, long v) { return __atomic_fetch_add((long *)p, v, __ATOMIC_RELAXED); }

#endif
