// rd_shim.cu - the L1 surface of the reference (oclhelper.h:12-39, helper.h:12-31) and the handful of raw OpenCL
// entry points its programs call (rect.cpp:60-64, poly.cpp:87-131, vidrect.cpp:120-127), on the CUDA runtime.
// No OpenCL runtime, no program build, no work-group autotuner: kernels are compiled ahead of time for sm_100a.
#include <stdarg.h>
#include <time.h>
#include <unistd.h>
#include <vector>
#include <mutex>
#include <ctype.h>
#include "rd_common.cuh"

std::atomic<int> g_rd_launches{0};

extern "C" {

// ---------------------------------------------------------------- helper.h
void exitf(int code, const char *mes, ...) {             // helper.c:31
  va_list ap;
  va_start(ap, mes);
  vfprintf(stderr, mes, ap);
  va_end(ap);
  fflush(stderr);
  exit(code);
}

int64_t currentTimeMillis(void) {                         // helper.c:105
  struct timespec ts;
  clock_gettime(CLOCK_REALTIME, &ts);
  return (int64_t)ts.tv_sec * 1000 + ts.tv_nsec / 1000000;
}

void sleepMillis(int ms) {                                // helper.c:115
  if (ms <= 0) return;
  struct timespec ts = {ms / 1000, (long)(ms % 1000) * 1000000L};
  nanosleep(&ts, NULL);
}

// uint64 -> void* map with 1024 buckets (helper.c:124-267).  Iteration order = bucket order, then insertion order
// with swap-remove; the host tail's output order depends on it, so the structure is kept.
struct ArrayMapNode { uint64_t key; void *value; };
struct ArrayMap { std::vector<ArrayMapNode> bucket[1024]; int total; };
static inline int am_hash(uint64_t key) { return (int)((key ^ (key >> 10) ^ (key >> 20) ^ (key >> 30)) & 1023); }

ArrayMap *initArrayMap(void) { ArrayMap *m = new ArrayMap(); m->total = 0; return m; }
void ArrayMap_dispose(ArrayMap *thiz) { delete thiz; }
int ArrayMap_size(ArrayMap *thiz) { return thiz->total; }
void *ArrayMap_remove(ArrayMap *thiz, uint64_t key) {
  std::vector<ArrayMapNode> &b = thiz->bucket[am_hash(key)];
  for (size_t i = 0; i < b.size(); i++)
    if (b[i].key == key) {
      void *old = b[i].value;
      b[i] = b.back();
      b.pop_back();
      thiz->total--;
      return old;
    }
  return NULL;
}
void *ArrayMap_put(ArrayMap *thiz, uint64_t key, void *value) {
  if (value == NULL) return ArrayMap_remove(thiz, key);
  std::vector<ArrayMapNode> &b = thiz->bucket[am_hash(key)];
  for (size_t i = 0; i < b.size(); i++)
    if (b[i].key == key) { void *old = b[i].value; b[i].value = value; return old; }
  ArrayMapNode n = {key, value};
  b.push_back(n);
  thiz->total++;
  return NULL;
}
void *ArrayMap_get(ArrayMap *thiz, uint64_t key) {
  std::vector<ArrayMapNode> &b = thiz->bucket[am_hash(key)];
  for (size_t i = 0; i < b.size(); i++) if (b[i].key == key) return b[i].value;
  return NULL;
}
uint64_t *ArrayMap_keyArray(ArrayMap *thiz) {
  uint64_t *a = (uint64_t *)malloc(sizeof(uint64_t) * (thiz->total > 0 ? thiz->total : 1));
  int p = 0;
  for (int j = 0; j < 1024; j++) for (size_t i = 0; i < thiz->bucket[j].size(); i++) a[p++] = thiz->bucket[j][i].key;
  return a;
}
void **ArrayMap_valueArray(ArrayMap *thiz) {
  void **a = (void **)malloc(sizeof(void *) * (thiz->total > 0 ? thiz->total : 1));
  int p = 0;
  for (int j = 0; j < 1024; j++) for (size_t i = 0; i < thiz->bucket[j].size(); i++) a[p++] = thiz->bucket[j][i].value;
  return a;
}

uint64_t ArrayMap_getKey(ArrayMap *thiz, int idx) {        // helper.h:30 (declared by the reference, defined nowhere in it): idx-th entry in
  for (int j = 0; j < 1024; j++) {                        // keyArray order
    if (idx < (int)thiz->bucket[j].size()) return thiz->bucket[j][idx].key;
    idx -= (int)thiz->bucket[j].size();
  }
  return 0;
}
void *ArrayMap_getValue(ArrayMap *thiz, int idx) {         // helper.h:31
  for (int j = 0; j < 1024; j++) {
    if (idx < (int)thiz->bucket[j].size()) return thiz->bucket[j][idx].value;
    idx -= (int)thiz->bucket[j].size();
  }
  return NULL;
}
// helper.c:40-62 : a whole text file as a malloc()ed string; larger than maxSize or unreadable ends the process
char *readFileAsStr(const char *fn, int maxSize) {
  FILE *fp = fopen(fn, "r");
  if (!fp) exitf(-1, "Couldn't open file %s\n", fn);
  fseek(fp, 0, SEEK_END);
  long size = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  if (size > maxSize) exitf(-1, "readFileAsStr : file too large (%d bytes)\n", (int)size);
  char *buf = (char *)malloc((size_t)size + 10);
  if (!buf) exitf(-1, "readFileAsStr : malloc failed\n");
  size = (long)fread(buf, 1, (size_t)size, fp);
  buf[size] = '\0';
  fclose(fp);
  return buf;
}
// helper.c:64-88 : the files of a NULL-terminated name list, concatenated (1 MB limit)
char *readFileAsStrN(const char **fn) {
  size_t total = 0;
  char *buf = (char *)malloc(10);
  buf[0] = '\0';
  for (int i = 0; fn[i] != NULL; i++) {
    char *one = readFileAsStr(fn[i], 1000000);
    const size_t len = strlen(one);
    if (total + len > 1000000) exitf(-1, "readFileAsStrN : total file size %d bytes is too large\n", (int)(total + len));
    buf = (char *)realloc(buf, total + len + 10);
    if (!buf) exitf(-1, "readFileAsStrN : realloc failed\n");
    memcpy(buf + total, one, len + 1);
    total += len;
    free(one);
  }
  return buf;
}
void String_trim(char *str) {                             // helper.c:90-101 : strip leading and trailing white space in place
  char *dst = str, *src = str, *end = str;
  while (*src != '\0' && isspace((unsigned char)*src)) src++;
  for (; *src != '\0'; src++) {
    *dst++ = *src;
    if (!isspace((unsigned char)*src)) end = dst;
  }
  *end = '\0';
}

// ---------------------------------------------------------------- oclhelper.h
const char *clStrError(int c) {                           // oclhelper.c:31-111 (only the codes this library can produce)
  switch (c) {
    case CL_SUCCESS: return "CL_SUCCESS";
    case CL_MEM_OBJECT_ALLOCATION_FAILURE: return "CL_MEM_OBJECT_ALLOCATION_FAILURE";
    case CL_OUT_OF_RESOURCES: return "CL_OUT_OF_RESOURCES";
    case CL_INVALID_VALUE: return "CL_INVALID_VALUE";
    case CL_INVALID_COMMAND_QUEUE: return "CL_INVALID_COMMAND_QUEUE";
    case CL_INVALID_MEM_OBJECT: return "CL_INVALID_MEM_OBJECT";
    default: return "Unknown error";
  }
}
cl_int checkError(cl_int ret, const char *s) {            // oclhelper.c:113-131
  if (ret != CL_SUCCESS) exitf(-1, "%s : %s\n", s ? s : "error", clStrError(ret));
  return ret;
}
cl_int ce(cl_int ret) { return checkError(ret, "Error"); } // oclhelper.c:133-138

static std::mutex g_dev_mutex;
static std::vector<rd_cl_device *> g_devices;
static int probe_devices() {
  std::lock_guard<std::mutex> lk(g_dev_mutex);
  if (!g_devices.empty()) return (int)g_devices.size();
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
  for (int i = 0; i < n; i++) { rd_cl_device *d = new rd_cl_device; d->ordinal = i; g_devices.push_back(d); }
  return n;
}
int rd_device_count(void) { return probe_devices(); }
const char *rd_version(void) { return "rectdetect_b200 0.1 (sm_100a)"; }
int rd_kernel_launches(void) { return g_rd_launches.load(); }

char *getDeviceName(cl_device_id device) {                // oclhelper.c:140-169 : whitespace -> '_'
  cudaDeviceProp p;
  RD_CUDA(cudaGetDeviceProperties(&p, device->ordinal));
  char *s = (char *)malloc(strlen(p.name) + 1);
  strcpy(s, p.name);
  for (char *c = s; *c; c++) if (*c == ' ' || *c == '\t') *c = '_';
  return s;
}
int simpleGetDevices(cl_device_id *devices, int maxDevices) {
  int n = probe_devices();
  if (n > maxDevices) n = maxDevices;
  for (int i = 0; i < n; i++) devices[i] = g_devices[i];
  return n;
}
cl_device_id simpleGetDevice(int did) {                   // oclhelper.c:171-197
  int n = probe_devices();
  if (n == 0) exitf(-1, "No platform available\n");     // no CUDA device: there is no CPU fallback
  if (did < 0 || did >= n) {
    if (did >= 0) fprintf(stderr, "Device %d does not exist\n", did);
    for (int i = 0; i < n; i++) { char *nm = getDeviceName(g_devices[i]); fprintf(stderr, "Device %d : %s\n", i, nm); free(nm); }
    exit(-1);
  }
  return g_devices[did];
}
cl_context simpleCreateContext(cl_device_id device) {     // oclhelper.c:225-233
  RD_CUDA(cudaSetDevice(device->ordinal));
  RD_CUDA(cudaFree(0));
  rd_cl_context *c = new rd_cl_context;
  c->ordinal = device->ordinal;
  return c;
}
// oclhelper.h:20-25 : the program-build / argument / launch helpers the reference's three host layers use internally.  There
// are no OpenCL program or kernel objects in this library (the kernels are sm_100a code linked into it, and no exported call
// creates a cl_program / cl_kernel), so these exist for link completeness and end the process if something calls them.
int simpleBuildProgram(cl_program, cl_device_id, const char *) {
  exitf(-1, "rectdetect_b200: simpleBuildProgram: there is no OpenCL program to build (kernels are precompiled CUDA)\n");
  return -1;
}
void simpleSetKernelArg(cl_kernel, const char *, ...) { exitf(-1, "rectdetect_b200: simpleSetKernelArg: no OpenCL kernel objects exist in this library\n"); }
cl_event runKernel1D(cl_command_queue, cl_kernel, int, size_t, int, ...) { exitf(-1, "rectdetect_b200: runKernel1D: no OpenCL kernel objects exist in this library\n"); return NULL; }
cl_event runKernel2D(cl_command_queue, cl_kernel, int, size_t, size_t, int, ...) { exitf(-1, "rectdetect_b200: runKernel2D: no OpenCL kernel objects exist in this library\n"); return NULL; }
cl_event runKernel1Dx(cl_command_queue, cl_kernel, int, size_t, const cl_event *) { exitf(-1, "rectdetect_b200: runKernel1Dx: no OpenCL kernel objects exist in this library\n"); return NULL; }
cl_event runKernel2Dx(cl_command_queue, cl_kernel, int, size_t, size_t, const cl_event *) { exitf(-1, "rectdetect_b200: runKernel2Dx: no OpenCL kernel objects exist in this library\n"); return NULL; }

void waitForEvent(cl_event ev) {                          // oclhelper.c:799-817, without the 15 ms poll (Q22)
  if (!ev) return;
  RD_CUDA(cudaEventSynchronize(ev->ev));
}
void clearPlan(void) {}
int loadPlan(const char *, cl_device_id) { return 0; }
void savePlan(const char *, cl_device_id) {}
void startProfiling(size_t, size_t, size_t) {}
void finishProfiling(void) {}
void showPlan(void) {}
static std::atomic<int> g_next_kid{0};
int getNextKernelID(void) { return g_next_kid.fetch_add(1); }

void *allocatePinnedMemory(size_t z, cl_context, cl_command_queue) {   // oclhelper.c:837-853
  void *p = NULL;
  RD_CUDA(cudaHostAlloc(&p, z ? z : 1, cudaHostAllocDefault));
  return p;
}
void freePinnedMemory(void *p, cl_context, cl_command_queue) { if (p) RD_CUDA(cudaFreeHost(p)); }

// ---------------------------------------------------------------- CL/cl.h compat
cl_command_queue clCreateCommandQueue(cl_context context, cl_device_id device, cl_command_queue_properties, cl_int *errcode_ret) {
  int ord = device ? device->ordinal : (context ? context->ordinal : 0);
  RD_CUDA(cudaSetDevice(ord));
  rd_cl_queue *q = new rd_cl_queue;
  q->ordinal = ord;
  q->owned = 1;
  RD_CUDA(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
  if (errcode_ret) *errcode_ret = CL_SUCCESS;
  return q;
}
cl_int clReleaseCommandQueue(cl_command_queue q) {
  if (!q) return CL_INVALID_COMMAND_QUEUE;
  if (q->owned) { RD_CUDA(cudaStreamSynchronize(q->stream)); RD_CUDA(cudaStreamDestroy(q->stream)); }
  delete q;
  return CL_SUCCESS;
}
cl_int clReleaseContext(cl_context c) { delete c; return CL_SUCCESS; }
cl_int clFlush(cl_command_queue) { return CL_SUCCESS; }
cl_int clFinish(cl_command_queue q) { RD_CUDA(cudaStreamSynchronize(rd_stream(q))); return CL_SUCCESS; }

cl_mem clCreateBuffer(cl_context context, cl_mem_flags flags, size_t size, void *host_ptr, cl_int *errcode_ret) {
  if (context) RD_CUDA(cudaSetDevice(context->ordinal));
  rd_cl_mem *m = new rd_cl_mem;
  m->bytes = size;
  m->owned = 1;
  cudaError_t e = cudaMalloc(&m->dptr, size ? size : 1);
  if (e != cudaSuccess) {
    cudaGetLastError();
    delete m;
    if (errcode_ret) { *errcode_ret = CL_MEM_OBJECT_ALLOCATION_FAILURE; return NULL; }
    exitf(-1, "clCreateBuffer : %s\n", cudaGetErrorString(e));
  }
  if ((flags & CL_MEM_COPY_HOST_PTR) && host_ptr) RD_CUDA(cudaMemcpy(m->dptr, host_ptr, size, cudaMemcpyHostToDevice));
  if (errcode_ret) *errcode_ret = CL_SUCCESS;
  return m;
}
cl_int clReleaseMemObject(cl_mem m) {
  if (!m) return CL_INVALID_MEM_OBJECT;
  if (m->owned) RD_CUDA(cudaFree(m->dptr));
  delete m;
  return CL_SUCCESS;
}
cl_int clRetainEvent(cl_event ev) { if (ev) ev->refs.fetch_add(1); return CL_SUCCESS; }
cl_int clReleaseEvent(cl_event ev) {
  if (ev && ev->refs.fetch_sub(1) == 1) { cudaEventDestroy(ev->ev); delete ev; }
  return CL_SUCCESS;
}
cl_int clEnqueueReadBuffer(cl_command_queue queue, cl_mem buffer, cl_bool blocking, size_t offset, size_t size, void *ptr,
                           cl_uint nwait, const cl_event *wait, cl_event *event) {
  cudaStream_t s = rd_stream(queue);
  for (cl_uint i = 0; i < nwait; i++) if (wait && wait[i]) RD_CUDA(cudaStreamWaitEvent(s, wait[i]->ev, 0));
  if (!buffer || offset + size > buffer->bytes) return CL_INVALID_VALUE;
  RD_CUDA(cudaMemcpyAsync(ptr, (char *)buffer->dptr + offset, size, cudaMemcpyDeviceToHost, s));
  if (event) { cl_event one = NULL; *event = rd_make_event(s, &one); }
  if (blocking) RD_CUDA(cudaStreamSynchronize(s));
  return CL_SUCCESS;
}
cl_int clEnqueueWriteBuffer(cl_command_queue queue, cl_mem buffer, cl_bool blocking, size_t offset, size_t size, const void *ptr,
                            cl_uint nwait, const cl_event *wait, cl_event *event) {
  cudaStream_t s = rd_stream(queue);
  for (cl_uint i = 0; i < nwait; i++) if (wait && wait[i]) RD_CUDA(cudaStreamWaitEvent(s, wait[i]->ev, 0));
  if (!buffer || offset + size > buffer->bytes) return CL_INVALID_VALUE;
  RD_CUDA(cudaMemcpyAsync((char *)buffer->dptr + offset, ptr, size, cudaMemcpyHostToDevice, s));
  if (event) { cl_event one = NULL; *event = rd_make_event(s, &one); }
  if (blocking) RD_CUDA(cudaStreamSynchronize(s));
  return CL_SUCCESS;
}

// ---------------------------------------------------------------- extensions
cl_mem rd_wrap_device_memory(void *dptr, size_t bytes) {
  rd_cl_mem *m = new rd_cl_mem;
  m->dptr = dptr; m->bytes = bytes; m->owned = 0;
  return m;
}
void *rd_mem_device_ptr(cl_mem mem) { return mem ? mem->dptr : NULL; }
size_t rd_mem_size(cl_mem mem) { return mem ? mem->bytes : 0; }
cl_command_queue rd_wrap_stream(void *cuda_stream, int device) {
  rd_cl_queue *q = new rd_cl_queue;
  q->stream = (cudaStream_t)cuda_stream; q->ordinal = device; q->owned = 0;
  return q;
}
void *rd_queue_stream(cl_command_queue queue) { return queue ? (void *)queue->stream : NULL; }
void rd_free(void *p) { free(p); }

}  // extern "C"

cl_event rd_make_event(cudaStream_t s, const cl_event *events) {
  if (events == NULL) return NULL;
  rd_cl_event *e = new rd_cl_event;
  e->refs.store(1);
  RD_CUDA(cudaEventCreateWithFlags(&e->ev, cudaEventDisableTiming));
  RD_CUDA(cudaEventRecord(e->ev, s));
  return e;
}

void rd_wait_events(cudaStream_t s, const cl_event *events) {
  if (events) for (int i = 0; events[i] != NULL; i++) RD_CUDA(cudaStreamWaitEvent(s, events[i]->ev, 0));
}

// ---------------------------------------------------------------- per-kernel timing (used by bench.py for the roofline line)
#include <map>
#include <string>
std::atomic<int> g_rd_prof_mode{0};
namespace {
struct ProfSlot { int name; cudaEvent_t a, b; };
std::mutex g_prof_mutex;
std::vector<ProfSlot> g_prof_slots;
std::vector<std::string> g_prof_names;
std::map<std::string, int> g_prof_index;
std::string g_prof_select;
std::string g_prof_report;
}

thread_local const char *g_rd_prof_stage = "";
void rd_prof_stage(const char *tag) { g_rd_prof_stage = tag ? tag : ""; }

int rd_prof_begin(const char *kname, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  if (g_rd_prof_mode.load() == 2 && strstr(kname, g_prof_select.c_str()) == NULL) return -1;
  // mode 3: the name carries the pipeline stage the calling thread is in ("A/kf_iir_h3"), for per-stage totals
  const std::string name = g_rd_prof_mode.load() == 3 ? std::string(g_rd_prof_stage) + "/" + kname : std::string(kname);
  auto it = g_prof_index.find(name);
  int idx;
  if (it == g_prof_index.end()) { idx = (int)g_prof_names.size(); g_prof_names.push_back(name); g_prof_index[name] = idx; }
  else idx = it->second;
  ProfSlot sl;
  sl.name = idx;
  RD_CUDA(cudaEventCreate(&sl.a));
  RD_CUDA(cudaEventCreate(&sl.b));
  RD_CUDA(cudaEventRecord(sl.a, s));
  g_prof_slots.push_back(sl);
  return (int)g_prof_slots.size() - 1;
}
void rd_prof_end(int slot, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  RD_CUDA(cudaEventRecord(g_prof_slots[slot].b, s));
}

extern "C" {
// mode 0 off, 1 all kernels, 2 kernels whose name contains `select`, 3 all kernels with the stage tag in the name
void rd_profile_start(int mode, const char *select) {
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  g_prof_select = select ? select : "";
  g_rd_prof_mode.store(mode);
}
// stops profiling, waits for the device and returns "name launches total_ms\n" lines (valid until the next call)
const char *rd_profile_stop(void) {
  g_rd_prof_mode.store(0);
  RD_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  std::vector<double> ms(g_prof_names.size(), 0.0);
  std::vector<int> cnt(g_prof_names.size(), 0);
  for (ProfSlot &sl : g_prof_slots) {
    float t = 0;
    if (cudaEventElapsedTime(&t, sl.a, sl.b) == cudaSuccess) { ms[sl.name] += t; cnt[sl.name]++; } else cudaGetLastError();
    cudaEventDestroy(sl.a);
    cudaEventDestroy(sl.b);
  }
  g_prof_slots.clear();
  g_prof_report.clear();
  char line[256];
  for (size_t i = 0; i < g_prof_names.size(); i++) {
    if (!cnt[i]) continue;
    snprintf(line, sizeof line, "%s %d %.6f\n", g_prof_names[i].c_str(), cnt[i], ms[i]);
    g_prof_report += line;
  }
  return g_prof_report.c_str();
}
}
