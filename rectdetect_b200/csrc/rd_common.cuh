// rd_common.cuh - shared definitions of librectdetect_b200.so (B200 / sm_100a only).
//
// The handle types the reference spells as OpenCL objects (cl_mem, cl_command_queue, ...) are
// plain structs around CUDA runtime objects.  Every kernel launch goes through RD_LAUNCH so that the library
// can report how many of its own kernels ran (rd_kernel_launches) and so that a launch failure ends the process
// the way the reference's ce()/checkError() do (oclhelper.c:113-138): message on stderr, exit(-1).
#ifndef RD_COMMON_CUH
#define RD_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <functional>
#include "../../include/rectdetect_b200.h"

struct rd_cl_platform { int unused; };
struct rd_cl_device   { int ordinal; };
struct rd_cl_context  { int ordinal; };
struct rd_cl_queue    { cudaStream_t stream; int ordinal; int owned; };
struct rd_cl_mem      { void *dptr; size_t bytes; int owned; };
struct rd_cl_event    { cudaEvent_t ev; std::atomic<int> refs; };
struct rd_cl_kernel   { int unused; };
struct rd_cl_program  { int unused; };

extern std::atomic<int> g_rd_launches;

#define RD_CUDA(call)                                                                                       \
  do {                                                                                                      \
    cudaError_t e_ = (call);                                                                                \
    if (e_ != cudaSuccess) exitf(-1, "rectdetect_b200: %s failed at %s:%d : %s\n", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
  } while (0)

// launch + count + check (cudaGetLastError only reports launch-configuration errors; execution errors
// surface at the next synchronising call, which is also RD_CUDA-checked)
#define RD_LAUNCH(kernel, grid, block, smem, stream, ...)                                                   \
  do {                                                                                                      \
    const int ps_ = g_rd_prof_mode.load(std::memory_order_relaxed) ? rd_prof_begin(#kernel, (stream)) : -1; \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                             \
    if (ps_ >= 0) rd_prof_end(ps_, (stream));                                                               \
    g_rd_launches.fetch_add(1, std::memory_order_relaxed);                                                  \
    cudaError_t e_ = cudaGetLastError();                                                                    \
    if (e_ != cudaSuccess) exitf(-1, "rectdetect_b200: launch of %s failed at %s:%d : %s\n", #kernel, __FILE__, __LINE__, cudaGetErrorString(e_)); \
  } while (0)

// per-kernel device timing with CUDA events on the launching stream (rd_profile_* in rectdetect_b200.h).
// mode 0: off (one relaxed load per launch), 1: every kernel, 2: only kernels whose name contains the selected string
extern std::atomic<int> g_rd_prof_mode;
int rd_prof_begin(const char *name, cudaStream_t s);
void rd_prof_end(int slot, cudaStream_t s);
void rd_prof_stage(const char *tag);      // stage tag of the calling thread (string literal), used by profile mode 3

// the stream behind a queue; also makes the queue's device current for the calling thread (a queue may belong to another device
// than the thread's current one: kernel launches and async allocations go to the current device)
static inline cudaStream_t rd_stream(cl_command_queue q) {
  if (!q) exitf(-1, "rectdetect_b200: NULL command queue\n");
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != q->ordinal) {
    if (cudaSetDevice(q->ordinal) != cudaSuccess) exitf(-1, "rectdetect_b200: cannot select device %d of the command queue\n", q->ordinal);
  }
  return q->stream;
}
template <typename T> static inline T *rd_ptr(cl_mem m) {
  if (!m) exitf(-1, "rectdetect_b200: NULL cl_mem\n");
  return (T *)m->dptr;
}
static inline void rd_need(cl_mem m, size_t bytes, const char *what) {
  if (!m || m->bytes < bytes) exitf(-1, "rectdetect_b200: buffer too small for %s (%zu < %zu bytes)\n", what, m ? m->bytes : (size_t)0, bytes);
}

// event returned by an operator: NULL when the caller passed events == NULL (oclhelper.c:740-743)
cl_event rd_make_event(cudaStream_t s, const cl_event *events);
// make `s` wait for every event of a NULL-terminated wait list (oclhelper.c:726-727)
void rd_wait_events(cudaStream_t s, const cl_event *events);

static inline int rd_cdiv(int a, int b) { return (a + b - 1) / b; }
static inline dim3 rd_grid2d(int iw, int ih, dim3 block) { return dim3(rd_cdiv(iw, block.x), rd_cdiv(ih, block.y)); }

// ---- frame batching -------------------------------------------------------------------------------------------
// Every kernel processes `nb` independent frames per launch.  All device buffers of one frame live in one arena and the
// arenas of a batch are `fs` bytes apart, so the pointers of frame z are the pointers of frame 0 plus z*fs, whatever
// they point at.  The frame index is blockIdx.z for 2-D kernels, blockIdx.y for 1-D kernels and blockIdx.x for the
// single-CTA-per-frame kernels.  The operator-level API (one caller-owned buffer set) launches with nb = 1, fs = 0.
static inline dim3 rd_gz(dim3 g, int nb) { g.z = nb; return g; }
static inline dim3 rd_gy(dim3 g, int nb) { g.y = nb; return g; }

// ---- device helpers shared by the three kernel families (oclimgutil.cl:28-63, identical copies in oclrect.cl:19-48) ----
#ifdef __CUDACC__
template <class... T> __device__ __forceinline__ void rd_batch_off(size_t o, T *&...p) { ((p = (T *)((char *)p + o)), ...); }
template <class... T> __device__ __forceinline__ void rd_batch_z(size_t fs, T *&...p) { rd_batch_off((size_t)blockIdx.z * fs, p...); }
template <class... T> __device__ __forceinline__ void rd_batch_y(size_t fs, T *&...p) { rd_batch_off((size_t)blockIdx.y * fs, p...); }
template <class... T> __device__ __forceinline__ void rd_batch_x(size_t fs, T *&...p) { rd_batch_off((size_t)blockIdx.x * fs, p...); }
__device__ __forceinline__ int rd_cl_clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
__device__ __forceinline__ int rd_mirror1(int x, int iw) { return rd_cl_clamp(x, -x, iw * 2 - 2 - x); }
__device__ __forceinline__ int rd_mirror(int x, int y, int iw, int ih) { return rd_mirror1(x, iw) + rd_mirror1(y, ih) * iw; }

// convert_uint_rtn + clamp.  Negative / NaN saturate to 0, which is what cvt.rmi.u32.f32 does.
__device__ __forceinline__ uint32_t rd_f2u_floor_sat(float v, uint32_t hi) { return min(__float2uint_rd(v), hi); }

__device__ __forceinline__ uint32_t rd_packlab(float l, float a, float b) {
  uint32_t ret = rd_f2u_floor_sat(__fmul_rn(b, 1024.0f), 1023u);
  ret = (ret << 10) | rd_f2u_floor_sat(__fmul_rn(a, 1024.0f), 1023u);
  ret = (ret << 12) | rd_f2u_floor_sat(__fmul_rn(l, 4096.0f), 4095u);
  return ret;
}
__device__ __forceinline__ void rd_unpacklab(uint32_t plab, float &l, float &a, float &b) {
  l = __fadd_rn(__fmul_rn((float)(int)(plab & 4095), 1.0f / 4096), 0.5f / 4096);
  a = __fadd_rn(__fmul_rn((float)(int)((plab >> 12) & 1023), 1.0f / 1024), 0.5f / 1024);
  b = __fadd_rn(__fmul_rn((float)(int)((plab >> 22) & 1023), 1.0f / 1024), 0.5f / 1024);
}
// canonical hypot / distance: exact double sum, IEEE double sqrt, one rounding to float
__device__ __forceinline__ float rd_hypot(float dx, float dy) {
  return (float)__dsqrt_rn(__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)));
}
__device__ __forceinline__ float rd_distance3(float dx, float dy, float dz) {
  return (float)__dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)), __dmul_rn((double)dz, (double)dz)));
}
__constant__ const int RD_RX[8] = {1, 1, 0, -1, -1, -1, 0, 1};   // oclrect.cl:12, oclpolyline.cl:63
__constant__ const int RD_RY[8] = {0, -1, -1, -1, 0, 1, 1, 1};

// shared-memory fetch-and-add issued by ONE elected lane.  (Written as PTX: for atomicAdd() the compiler emits its own
// warp-aggregation sequence - match, ballot, leader election, ~30 instructions - around every call, which is pure overhead
// when the caller has already aggregated the warp's contribution.)
__device__ __forceinline__ int rd_smem_fetch_add(int *p, int v) {
  int old;
  asm volatile("atom.shared.add.s32 %0, [%1], %2;" : "=r"(old) : "r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
  return old;
}

// ---- union-find with "the smaller index is the root" : every CCL of the path is built on it ----
__device__ __forceinline__ int rd_uf_find(const int *L, int x) {
  int p = __ldcg(L + x);          // L2 reads: parents are updated by atomics from other SMs
  while (p != x) { x = p; p = __ldcg(L + x); }
  return x;
}
__device__ __forceinline__ void rd_uf_unite(int *L, int a, int b) {
  for (;;) {
    a = rd_uf_find(L, a);
    b = rd_uf_find(L, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }      // a > b : hang a under b
    int old = atomicMin(L + a, b);
    if (old == a) return;
    a = old;                                     // somebody re-parented a meanwhile; retry from there
  }
}
#endif

// kernel families (host-side entry points used across translation units)
// CCL: label = smallest linear index of the 8-connected equal-value component, bgc pixels -> -1
void rd_label8x(int *label, const int *pix, void *scratch /* iw*ih bytes */, int bgc, int iw, int ih, int nb, size_t fs, cudaStream_t s);

// host tail (rd_tail.cpp) on the compact read-back record (see rd_rect.cu : tail_gather)
struct rd_tail_sample { int32_t segid; int32_t vote[5]; };            // vote = the table entry the (ls,segid) pair hashes to
#define RD_TAIL_NSAMPLE 15
typedef void (*rd_parallel_for_t)(int n, const std::function<void(int)> &fn);     // fn(0) .. fn(n-1), returns when all are done
rect_t *rd_tail_compact(const linesegment_t *ls /* n+1 entries */, const rd_tail_sample *samples /* (n+1)*15 */, int iw, int ih, double tanAOV,
                        rd_parallel_for_t pfor /* NULL: serial */);

#endif
