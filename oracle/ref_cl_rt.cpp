// ref_cl_rt.cpp - a synchronous single-device host "OpenCL runtime", just large enough for the REFERENCE's own host code
// (oclhelper.c, oclimgutil.c, oclpolyline.c, oclrect.c, compiled unmodified from /root/reference) to run its schedules
// over the reference's own kernels (the three .cl files compiled as C++ by cl_translate.py + cl_compat.h).
//
// TEST INFRASTRUCTURE: part of oracle/_ref/librd_ref.so (oracle/Makefile target _ref).  Never linked into the product.
//
// Execution model: every enqueue runs to completion before it returns (an in-order queue with no overlap).  An NDRange
// is executed row by row; with rd_ref_set_threads(1) (the default) the work-items run in raster order on the calling
// thread - ONE legal schedule of the kernels, deterministic, used to pin the oracle - and with n > 1 the rows are spread
// over n OpenMP threads the way an OpenCL CPU device spreads work-groups over cores (atomics are real atomics, the
// reference's in-place races are live): that mode is the timed CPU baseline (bench.py --impl reference).
#include <CL/cl.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

thread_local int rd_cl_gid[3] = {0, 0, 0};

typedef void (*rd_cl_tramp)(void **args, long gw, long gh, int nthreads);
struct rd_cl_kdef { const char *tag, *name; rd_cl_tramp fn; int nargs; unsigned ptrmask; };
static rd_cl_kdef g_kdefs[256];
static int g_nkdefs = 0;
static int g_threads = 1;
static long g_launches = 0, g_limit = -1;
static const char *g_trace[1 << 16];

extern "C" void rd_cl_register(const char *tag, const char *name, rd_cl_tramp fn, int nargs, unsigned ptrmask) {
  if (g_nkdefs >= 256) abort();
  g_kdefs[g_nkdefs++] = rd_cl_kdef{tag, name, fn, nargs, ptrmask};
}

struct rd_cl_platform { int unused; };
struct rd_cl_device { int unused; };
struct rd_cl_context { int unused; };
struct rd_cl_queue { int unused; };
struct rd_cl_mem { void *ptr; size_t bytes; };
struct rd_cl_event { int refs; };
struct rd_cl_program { char tag[32]; };
struct rd_cl_kernel { const rd_cl_kdef *def; unsigned long long val[24]; void *argp[24]; unsigned setmask; };

static rd_cl_platform g_platform;
static rd_cl_device g_device;

static cl_int info(const char *s, size_t size, void *value, size_t *size_ret) {
  const size_t n = strlen(s) + 1;
  if (size_ret) *size_ret = n;
  if (value) { if (size < n) { if (size) { memcpy(value, s, size - 1); ((char *)value)[size - 1] = 0; } } else memcpy(value, s, n); }
  return CL_SUCCESS;
}
static void done(cl_event *event) { if (event) { *event = new rd_cl_event{1}; } }

extern "C" {
void rd_ref_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int rd_ref_get_threads(void) { return g_threads; }
// launch trace / launch limit: lets a test stop the reference's schedule after its n-th kernel launch and look at the planes
long rd_ref_launches(void) { return g_launches; }
void rd_ref_trace_reset(void) { g_launches = 0; }
const char *rd_ref_trace_name(long i) { return i >= 0 && i < g_launches && i < (1 << 16) ? g_trace[i] : ""; }
void rd_ref_set_launch_limit(long n) { g_limit = n; }
void *rd_ref_mem_ptr(cl_mem m) { return m ? m->ptr : NULL; }
size_t rd_ref_mem_bytes(cl_mem m) { return m ? m->bytes : 0; }

cl_int clGetPlatformIDs(cl_uint n, cl_platform_id *platforms, cl_uint *nret) {
  if (platforms && n) platforms[0] = &g_platform;
  if (nret) *nret = 1;
  return CL_SUCCESS;
}
cl_int clGetDeviceIDs(cl_platform_id, cl_device_type, cl_uint n, cl_device_id *devices, cl_uint *nret) {
  if (devices && n) devices[0] = &g_device;
  if (nret) *nret = 1;
  return CL_SUCCESS;
}
cl_int clGetDeviceInfo(cl_device_id, cl_device_info what, size_t size, void *value, size_t *size_ret) {
  switch (what) {
  case CL_DEVICE_NAME: return info("host cores (reference .cl kernels compiled as C++)", size, value, size_ret);
  case CL_DEVICE_VERSION: return info("OpenCL 1.2 rd_ref", size, value, size_ret);
  case CL_DEVICE_EXTENSIONS: return info("", size, value, size_ret);   // no 64-bit atomics extension: the kernels' emulated xatom_add path
  }
  return CL_INVALID_VALUE;
}
cl_context clCreateContext(const cl_context_properties *, cl_uint, const cl_device_id *, void (*)(const char *, const void *, size_t, void *), void *, cl_int *err) {
  if (err) *err = CL_SUCCESS;
  return new rd_cl_context{0};
}
cl_int clReleaseContext(cl_context c) { delete c; return CL_SUCCESS; }
cl_command_queue clCreateCommandQueue(cl_context, cl_device_id, cl_command_queue_properties, cl_int *err) {
  if (err) *err = CL_SUCCESS;
  return new rd_cl_queue{0};
}
cl_int clReleaseCommandQueue(cl_command_queue q) { delete q; return CL_SUCCESS; }
cl_int clFlush(cl_command_queue) { return CL_SUCCESS; }
cl_int clFinish(cl_command_queue) { return CL_SUCCESS; }

cl_mem clCreateBuffer(cl_context, cl_mem_flags flags, size_t size, void *host_ptr, cl_int *err) {
  rd_cl_mem *m = new rd_cl_mem{NULL, size};
  if (posix_memalign(&m->ptr, 256, size ? size : 1) != 0) { delete m; if (err) *err = CL_MEM_OBJECT_ALLOCATION_FAILURE; return NULL; }
  // the contents of a fresh OpenCL buffer are undefined; zero-filled here (what fresh pages give a CPU device), which is also
  // the canonical choice for the reference's read-before-write planes
  memset(m->ptr, 0, size);
  if ((flags & CL_MEM_COPY_HOST_PTR) && host_ptr) memcpy(m->ptr, host_ptr, size);
  if (err) *err = CL_SUCCESS;
  return m;
}
cl_int clReleaseMemObject(cl_mem m) { if (m) { free(m->ptr); delete m; } return CL_SUCCESS; }
cl_int clReleaseEvent(cl_event e) { if (e && --e->refs == 0) delete e; return CL_SUCCESS; }
cl_int clRetainEvent(cl_event e) { if (e) e->refs++; return CL_SUCCESS; }
cl_int clGetEventInfo(cl_event, cl_event_info what, size_t size, void *value, size_t *size_ret) {
  if (what != CL_EVENT_COMMAND_EXECUTION_STATUS || size < sizeof(cl_int)) return CL_INVALID_VALUE;
  *(cl_int *)value = CL_COMPLETE;
  if (size_ret) *size_ret = sizeof(cl_int);
  return CL_SUCCESS;
}
cl_int clGetEventProfilingInfo(cl_event, cl_profiling_info, size_t size, void *value, size_t *size_ret) {
  if (size < sizeof(cl_ulong)) return CL_INVALID_VALUE;
  *(cl_ulong *)value = 0;
  if (size_ret) *size_ret = sizeof(cl_ulong);
  return CL_SUCCESS;
}
cl_int clEnqueueReadBuffer(cl_command_queue, cl_mem m, cl_bool, size_t offset, size_t size, void *ptr, cl_uint, const cl_event *, cl_event *event) {
  if (!m || offset + size > m->bytes) return CL_INVALID_VALUE;
  memcpy(ptr, (char *)m->ptr + offset, size);
  done(event);
  return CL_SUCCESS;
}
cl_int clEnqueueWriteBuffer(cl_command_queue, cl_mem m, cl_bool, size_t offset, size_t size, const void *ptr, cl_uint, const cl_event *, cl_event *event) {
  if (!m || offset + size > m->bytes) return CL_INVALID_VALUE;
  memcpy((char *)m->ptr + offset, ptr, size);
  done(event);
  return CL_SUCCESS;
}
void *clEnqueueMapBuffer(cl_command_queue, cl_mem m, cl_bool, cl_map_flags, size_t offset, size_t, cl_uint, const cl_event *, cl_event *event, cl_int *err) {
  if (err) *err = CL_SUCCESS;
  done(event);
  return (char *)m->ptr + offset;
}
cl_int clEnqueueUnmapMemObject(cl_command_queue, cl_mem, void *, cl_uint, const cl_event *, cl_event *event) { done(event); return CL_SUCCESS; }

cl_program clCreateProgramWithSource(cl_context, cl_uint count, const char **strings, const size_t *, cl_int *err) {
  rd_cl_program *p = new rd_cl_program;
  snprintf(p->tag, sizeof(p->tag), "%s", count ? strings[0] : "");
  if (err) *err = CL_SUCCESS;
  return p;
}
cl_int clBuildProgram(cl_program, cl_uint, const cl_device_id *, const char *, void (*)(cl_program, void *), void *) { return CL_SUCCESS; }
cl_int clGetProgramBuildInfo(cl_program, cl_device_id, cl_program_build_info, size_t size, void *value, size_t *size_ret) { return info("", size, value, size_ret); }
cl_int clReleaseProgram(cl_program p) { delete p; return CL_SUCCESS; }
cl_kernel clCreateKernel(cl_program p, const char *name, cl_int *err) {
  for (int i = 0; i < g_nkdefs; i++)
    if (!strcmp(g_kdefs[i].tag, p->tag) && !strcmp(g_kdefs[i].name, name)) {
      rd_cl_kernel *k = new rd_cl_kernel;
      memset(k, 0, sizeof(*k));
      k->def = &g_kdefs[i];
      if (err) *err = CL_SUCCESS;
      return k;
    }
  if (err) *err = CL_INVALID_KERNEL_NAME;
  return NULL;   // as a real runtime: the reference never checks (a few of its kernel names do not exist in the .cl files)
}
cl_int clReleaseKernel(cl_kernel k) { delete k; return CL_SUCCESS; }
cl_int clGetKernelInfo(cl_kernel k, cl_kernel_info what, size_t size, void *value, size_t *size_ret) {
  if (!k || what != CL_KERNEL_FUNCTION_NAME) return CL_INVALID_VALUE;
  return info(k->def->name, size, value, size_ret);
}
cl_int clSetKernelArg(cl_kernel k, cl_uint index, size_t size, const void *value) {
  if (!k) return CL_INVALID_VALUE;
  if ((int)index >= k->def->nargs) return CL_INVALID_ARG_INDEX;
  if (size > 8) return CL_INVALID_ARG_SIZE;
  if (k->def->ptrmask >> index & 1) {
    if (size != sizeof(cl_mem)) return CL_INVALID_ARG_SIZE;
    cl_mem m = *(const cl_mem *)value;
    void *p = m ? m->ptr : NULL;
    memcpy(&k->val[index], &p, sizeof(p));
  } else {
    k->val[index] = 0;
    memcpy(&k->val[index], value, size);
  }
  k->argp[index] = &k->val[index];
  k->setmask |= 1u << index;
  return CL_SUCCESS;
}
cl_int clEnqueueNDRangeKernel(cl_command_queue, cl_kernel k, cl_uint dim, const size_t *, const size_t *gws, const size_t *, cl_uint, const cl_event *, cl_event *event) {
  if (!k) return CL_INVALID_VALUE;
  if (k->setmask != (k->def->nargs >= 32 ? ~0u : (1u << k->def->nargs) - 1)) { fprintf(stderr, "rd_ref: kernel %s launched with unset arguments\n", k->def->name); abort(); }
  if (g_limit < 0 || g_launches < g_limit) k->def->fn(k->argp, (long)gws[0], dim > 1 ? (long)gws[1] : 1, g_threads);
  if (g_launches < (1 << 16)) g_trace[g_launches] = k->def->name;
  g_launches++;
  done(event);
  return CL_SUCCESS;
}
}  // extern "C"
