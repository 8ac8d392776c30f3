/* rectdetect_b200.h - the C ABI of librectdetect_b200.so.
 *
 * The library is a drop-in for the object files the reference links into its programs (oclhelper.o,
 * helper.o, oclimgutil.o, oclpolyline.o, oclrect.o + libOpenCL): it exports the SAME C symbols with the same
 * argument meaning, backed by hand-written sm_100a CUDA kernels instead of OpenCL C compiled at run time.
 * Every declaration below cites the reference declaration it replaces (file:line under /root/reference).
 * A program written against the reference's own headers keeps compiling: it only needs include/CL/cl.h from
 * this tree in place of the Khronos header.
 *
 * Error convention (reference: helper.c:31, oclhelper.c:113-138): there are no error codes - any failure
 * prints a message to stderr and calls exit(-1).  In particular the library exits if no CUDA device is
 * usable: there is no CPU fallback.
 */
#ifndef RECTDETECT_B200_H
#define RECTDETECT_B200_H

#include <stddef.h>
#include <stdint.h>
#include "CL/cl.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------------
 * Plain-data types that cross the boundary
 * ---------------------------------------------------------------------------------------------------- */
#ifndef RD_HAVE_REFERENCE_HEADERS   /* define this when vec234.h / oclpolyline.h / oclrect.h of the reference are included too */
typedef struct { double a[2]; } vec2;                     /* vec234.h:5-7  */
typedef struct { double a[3]; } vec3;                     /* vec234.h:64-66 */

typedef struct linesegment_t {                            /* oclpolyline.h:74-83, 56 bytes; element 0 of a list is a header
                                                             whose first int is the number of segments n, valid ids 1..n */
  float x0, y0, x1, y1;
  int32_t startIndex, endIndex;
  int32_t leftPtr, rightPtr;                              /* chain links, 0 = none */
  int32_t startCount, endCount;
  int32_t maxDist;
  int32_t polyid;                                         /* 0 = dead entry */
  int32_t npix;
  int32_t level;
} linesegment_t;

typedef struct rect_t {                                   /* oclrect.h:5-15, 176 bytes; element 0 is a header (nItems) */
  union {
    struct {
      vec2 c2[4];                                         /* image-plane corners */
      vec3 c3[4];                                         /* estimated 3-D corners */
      double value;                                       /* pose residual */
      uint32_t status;                                    /* bit0: looks like a screen, bit1: from a polyline chain */
    };
    int nItems;
  };
} rect_t;

typedef struct oclimgutil_t oclimgutil_t;                 /* oclimgutil.h:5-72  (opaque here) */
typedef struct oclpolyline_t oclpolyline_t;               /* oclpolyline.h:5-72 (opaque here) */
#endif
struct oclrect_t;                                         /* oclrect.c:41-98 */
typedef struct ArrayMap ArrayMap;                         /* helper.h:21 */

/* ------------------------------------------------------------------------------------------------------
 * L1 - runtime helper surface the programs use (oclhelper.h:12-39, helper.h:12-31)
 * ---------------------------------------------------------------------------------------------------- */
const char *clStrError(int c);                            /* oclhelper.h:12 */
cl_int checkError(cl_int ret, const char *s);             /* oclhelper.h:13 : non-success -> message + exit(-1) */
cl_int ce(cl_int ret);                                    /* oclhelper.h:14 */
char *getDeviceName(cl_device_id device);                 /* oclhelper.h:16 : malloc()ed string */
cl_device_id simpleGetDevice(int did);                    /* oclhelper.h:17 : did < 0 lists devices and exits */
int simpleGetDevices(cl_device_id *devices, int maxDevices);   /* oclhelper.h:18 */
cl_context simpleCreateContext(cl_device_id device);      /* oclhelper.h:19 */
int simpleBuildProgram(cl_program program, cl_device_id device, const char *optionString);           /* oclhelper.h:20 : link stubs - no OpenCL   */
void simpleSetKernelArg(cl_kernel kernel, const char *format, ...);                                   /* oclhelper.h:21   program / kernel objects  */
cl_event runKernel1D(cl_command_queue queue, cl_kernel kernel, int kernelID, size_t ws1, int nev, ...);            /* :22   exist in this library;   */
cl_event runKernel2D(cl_command_queue queue, cl_kernel kernel, int kernelID, size_t ws1, size_t ws2, int nev, ...); /* :23  calling one ends the     */
cl_event runKernel1Dx(cl_command_queue queue, cl_kernel kernel, int kernelID, size_t ws1, const cl_event *events);  /* :24  process                  */
cl_event runKernel2Dx(cl_command_queue queue, cl_kernel kernel, int kernelID, size_t ws1, size_t ws2, const cl_event *events); /* :25 */
void waitForEvent(cl_event ev);                           /* oclhelper.h:27 : cudaEventSynchronize, no 15 ms polling */
void clearPlan(void);                                     /* oclhelper.h:29-34 : the work-group autotuner has no CUDA */
int loadPlan(const char *fn, cl_device_id device);        /*   counterpart; loadPlan returns 0 ("plan found") so   */
void savePlan(const char *fn, cl_device_id device);       /*   rect.cpp:86 skips its 48-run sweep, the rest are     */
void startProfiling(size_t ws1, size_t ws2, size_t ws3);  /*   no-ops.                                              */
void finishProfiling(void);
void showPlan(void);
void *allocatePinnedMemory(size_t z, cl_context context, cl_command_queue queue);   /* oclhelper.h:36 : cudaHostAlloc */
void freePinnedMemory(void *p, cl_context context, cl_command_queue queue);         /* oclhelper.h:37 */
int getNextKernelID(void);                                /* oclhelper.h:39 */

void exitf(int code, const char *mes, ...);               /* helper.h:12 */
char *readFileAsStr(const char *fn, int maxSize);         /* helper.h:13 : malloc()ed contents */
char *readFileAsStrN(const char **fn);                    /* helper.h:14 : NULL-terminated list of files, concatenated */
void String_trim(char *str);                              /* helper.h:17 */
int64_t currentTimeMillis(void);                          /* helper.h:15 */
void sleepMillis(int ms);                                 /* helper.h:16 */
ArrayMap *initArrayMap(void);                             /* helper.h:23-31 : uint64 -> void* map, 1024 buckets */
void ArrayMap_dispose(ArrayMap *thiz);
int ArrayMap_size(ArrayMap *thiz);
void *ArrayMap_remove(ArrayMap *thiz, uint64_t key);
void *ArrayMap_put(ArrayMap *thiz, uint64_t key, void *value);
void *ArrayMap_get(ArrayMap *thiz, uint64_t key);
uint64_t *ArrayMap_keyArray(ArrayMap *thiz);
void **ArrayMap_valueArray(ArrayMap *thiz);
uint64_t ArrayMap_getKey(ArrayMap *thiz, int idx);        /* helper.h:30-31 : declared by the reference but defined nowhere in it; here: the idx-th */
void *ArrayMap_getValue(ArrayMap *thiz, int idx);         /*                  entry in keyArray order */

/* ------------------------------------------------------------------------------------------------------
 * L2 - image operators (oclimgutil.h:74-100).  Every operator is an asynchronous enqueue on `queue`'s stream.
 * The returned cl_event is NULL whenever `events` is NULL (all call sites of the reference); otherwise it is an
 * event the caller releases with clReleaseEvent.  Buffers are planes of iw*ih 32-bit words unless noted.
 * ---------------------------------------------------------------------------------------------------- */
oclimgutil_t *init_oclimgutil(cl_device_id device, cl_context context);                                             /* oclimgutil.h:74 */
void dispose_oclimgutil(oclimgutil_t *thiz);                                                                        /* oclimgutil.h:75 */
cl_event oclimgutil_clear(oclimgutil_t *thiz, cl_mem out, int size /*bytes*/, cl_command_queue queue, const cl_event *events);          /* :77 */
cl_event oclimgutil_copy(oclimgutil_t *thiz, cl_mem out, cl_mem in, int size /*bytes*/, cl_command_queue queue, const cl_event *events); /* :78 */
cl_event oclimgutil_cast_i_f(oclimgutil_t *thiz, cl_mem out, cl_mem in, float scale, int size, cl_command_queue queue, const cl_event *events); /* :79 */
cl_event oclimgutil_cast_c_i(oclimgutil_t *thiz, cl_mem out, cl_mem in, int size, cl_command_queue queue, const cl_event *events);      /* :80 */
cl_event oclimgutil_threshold_i_i(oclimgutil_t *thiz, cl_mem out, cl_mem in, int vlow, int threshold, int vhigh, int size, cl_command_queue queue, const cl_event *events); /* :81 */
cl_event oclimgutil_threshold_f_f(oclimgutil_t *thiz, cl_mem out, cl_mem in, float vlow, float threshold, float vhigh, int size, cl_command_queue queue, const cl_event *event); /* :82 */
cl_event oclimgutil_rand(oclimgutil_t *thiz, cl_mem out, int size, cl_command_queue queue, const cl_event *events);                     /* :83 (seed 0) */
cl_event oclimgutil_convert_bgr_luminancef(oclimgutil_t *thiz, cl_mem out, cl_mem in, int iw, int ih, int ws, cl_command_queue queue, const cl_event *events); /* :84 */
cl_event oclimgutil_convert_bgr_lumaf(oclimgutil_t *thiz, cl_mem out, cl_mem in, float f, int iw, int ih, int ws, cl_command_queue queue, const cl_event *events); /* :85 */
cl_event oclimgutil_convert_bgr_labeli(oclimgutil_t *thiz, cl_mem out, cl_mem in, int bgc, int iw, int ih, int ws, cl_command_queue queue, const cl_event *events); /* :86 */
cl_event oclimgutil_edge_f_f(oclimgutil_t *thiz, cl_mem out, cl_mem in, int iw, int ih, cl_command_queue queue, const cl_event *events); /* :87 */
cl_event oclimgutil_edgevec_f2_f(oclimgutil_t *thiz, cl_mem out /*2 floats/px*/, cl_mem in, int iw, int ih, cl_command_queue queue, const cl_event *events); /* :88 */
cl_event oclimgutil_thinthres_f_f_f2(oclimgutil_t *thiz, cl_mem out, cl_mem in, cl_mem vec, int iw, int ih, cl_command_queue queue, const cl_event *events); /* :89 */
cl_event oclimgutil_thincubic_f_f_f2(oclimgutil_t *thiz, cl_mem out, cl_mem in, cl_mem vec, int iw, int ih, cl_command_queue queue, const cl_event *events); /* :90 */
cl_event oclimgutil_label8x_int_int(oclimgutil_t *thiz, cl_mem out, cl_mem in, cl_mem tmp, int bgc, int iw, int ih, cl_command_queue queue, const cl_event *events); /* :91 */
cl_event oclimgutil_iirblur_f_f(oclimgutil_t *thiz, cl_mem obuf, cl_mem ibuf, cl_mem tmp0, cl_mem tmp1, int r, int iw, int ih, cl_command_queue queue, const cl_event *events); /* :92 */
cl_event oclimgutil_convert_plab_bgr(oclimgutil_t *thiz, cl_mem out, cl_mem in, int iw, int ih, int ws, cl_command_queue queue, const cl_event *events); /* :93 - BGR8 -> packed Lab (names are swapped in the reference) */
cl_event oclimgutil_convert_bgr_plab(oclimgutil_t *thiz, cl_mem out, cl_mem in, int iw, int ih, int ws, cl_command_queue queue, const cl_event *events); /* :94 - packed Lab -> BGR8 */
cl_event oclimgutil_unpack_f_f_f_plab(oclimgutil_t *thiz, cl_mem out0, cl_mem out1, cl_mem out2, cl_mem in, int iw, int ih, cl_command_queue queue, const cl_event *events); /* :95 */
cl_event oclimgutil_pack_plab_f_f_f(oclimgutil_t *thiz, cl_mem out, cl_mem in0, cl_mem in1, cl_mem in2, int iw, int ih, cl_command_queue queue, const cl_event *events); /* :96 */
cl_event oclimgutil_edgevec_f2_plab(oclimgutil_t *thiz, cl_mem out, cl_mem in, int iw, int ih, cl_command_queue queue, const cl_event *events); /* :97 */
cl_event oclimgutil_edge_f_plab(oclimgutil_t *thiz, cl_mem out, cl_mem in, int iw, int ih, cl_command_queue queue, const cl_event *events); /* :98 */
cl_event oclimgutil_calcStrength(oclimgutil_t *thiz, cl_mem out, cl_mem edge, cl_mem label, int iw, int ih, cl_command_queue queue, const cl_event *events); /* :99 */
cl_event oclimgutil_filterStrength(oclimgutil_t *thiz, cl_mem labelinout, cl_mem str, int thre, int iw, int ih, cl_command_queue queue, const cl_event *events); /* :100 */

/* ------------------------------------------------------------------------------------------------------
 * L2 - polyline (oclpolyline.h:85-88).  lsList: lsListSize bytes (callers pass iw*ih*16); tmp0 is the big scratch
 * (>= iw*ih*16 bytes), tmp1..tmp6 are planes.  in: 0/1 edge mask.  Outputs: lsIdOut (per-pixel segment id), lsList.
 * ---------------------------------------------------------------------------------------------------- */
oclpolyline_t *init_oclpolyline(cl_device_id device, cl_context context);                                           /* oclpolyline.h:85 */
void dispose_oclpolyline(oclpolyline_t *thiz);                                                                      /* oclpolyline.h:86 */
cl_event oclpolyline_execute(oclpolyline_t *thiz, cl_mem lsList, int lsListSize, cl_mem lsIdOut, cl_mem in, cl_mem tmp0, cl_mem tmp1, cl_mem tmp2,
                             cl_mem tmp3, cl_mem tmp4, cl_mem tmp5, cl_mem tmp6, float minerror, int sizeThre, int iw, int ih,
                             cl_command_queue queue, const cl_event *events);                                       /* oclpolyline.h:88 */

/* ------------------------------------------------------------------------------------------------------
 * L3 - the rectangle detector object (oclrect.h:17-23)
 * ---------------------------------------------------------------------------------------------------- */
struct oclrect_t *init_oclrect(oclimgutil_t *oclimgutil, oclpolyline_t *oclpolyline, cl_device_id device, cl_context context,
                               cl_command_queue queue, int iw, int ih);                                             /* oclrect.h:17 */
void dispose_oclrect(struct oclrect_t *thiz);                                                                       /* oclrect.h:18 */
/* synchronous: BGR8 frame (row stride ws) in host memory -> malloc()ed rect_t list, element 0 = header (oclrect.h:20) */
rect_t *oclrect_executeOnce(struct oclrect_t *thiz, uint8_t *imgData, int ws, const double tanAOV);
/* two-deep pipeline: enqueue frame N+1 before polling frame N (oclrect.h:22-23, vidrect.cpp:159-172) */
void oclrect_enqueueTask(struct oclrect_t *thiz, uint8_t *imgData, int ws);
rect_t *oclrect_pollTask(struct oclrect_t *thiz, const double tanAOV);

/* ------------------------------------------------------------------------------------------------------
 * Extensions (no counterpart in the reference): adopting existing device memory / streams, the Stage-B/D kernels
 * of oclrect.cl as individually callable operators (for operator-level parity tests), introspection and the
 * frame-batch engine used for throughput (frames are independent).
 * ---------------------------------------------------------------------------------------------------- */
cl_mem rd_wrap_device_memory(void *dptr, size_t bytes);          /* non-owning cl_mem around memory you allocated (e.g. a torch tensor) */
void  *rd_mem_device_ptr(cl_mem mem);
size_t rd_mem_size(cl_mem mem);
cl_command_queue rd_wrap_stream(void *cuda_stream, int device);  /* non-owning queue around an existing cudaStream_t */
void  *rd_queue_stream(cl_command_queue queue);
int    rd_device_count(void);                                    /* 0 when no CUDA device/driver is usable; never exits */
const char *rd_version(void);
int    rd_kernel_launches(void);                                 /* kernels launched by this library so far (process-wide counter) */
/* per-kernel device time from CUDA events recorded on the launching stream around each launch.
 * mode 1: every kernel, mode 2: only kernels whose name contains `select`, mode 3: every kernel, names prefixed with the
 * stage of the production schedule ("A/", "B/", "C/", "D/").  rd_profile_stop waits for the device and
 * returns one "kernel_name launches total_ms" line per kernel (string owned by the library, valid until the next call). */
void   rd_profile_start(int mode, const char *select);
const char *rd_profile_stop(void);

/* Stage B / D operators, 1:1 with the __kernel functions of oclrect.cl (line numbers in parentheses) */
void rd_rect_simpleJunction(cl_mem out, cl_mem in, int iw, int ih, cl_command_queue q);                    /* oclrect.cl:74  */
void rd_rect_simpleConnect(cl_mem out, cl_mem in, int iw, int ih, cl_command_queue q);                     /* oclrect.cl:97  */
void rd_rect_stringify(cl_mem out, cl_mem in, int mod2, int iw, int ih, cl_command_queue q);               /* oclrect.cl:123 */
void rd_rect_blblur0(cl_mem out, cl_mem edge_i8, cl_mem in, int iw, int ih, cl_command_queue q);           /* oclrect.cl:155 */
void rd_rect_blblur1(cl_mem out, cl_mem edge_i8, cl_mem in, int iw, int ih, cl_command_queue q);           /* oclrect.cl:181 */
void rd_rect_quantize(cl_mem out, cl_mem in, int n0, int n1, int n2, int iw, int ih, cl_command_queue q);  /* oclrect.cl:207 */
void rd_rect_despeckle(cl_mem out, cl_mem in, cl_mem edge, int iw, int ih, cl_command_queue q);            /* oclrect.cl:218 */
void rd_rect_mkMergeMask0(cl_mem out, cl_mem junction, int iw, int ih, cl_command_queue q);                /* oclrect.cl:246 */
void rd_rect_mkMergeMask1(cl_mem inout, cl_mem junction, int iw, int ih, cl_command_queue q);              /* oclrect.cl:263 */
void rd_rect_labelMerge(cl_mem label, cl_mem pix, cl_mem mask, cl_mem edge, int iw, int ih, cl_command_queue q); /* oclrect.cl:289-334 + oclrect.c:325-331, converged */
void rd_rect_calcSize(cl_mem out, cl_mem label, int iw, int ih, cl_command_queue q);                       /* oclrect.cl:336 */
void rd_rect_despeckle2(cl_mem labelinout, cl_mem size, cl_mem scratch, int thre, int iw, int ih, cl_command_queue q); /* oclrect.cl:348, in place, work-items in raster order */
void rd_rect_markBoundary(cl_mem out, cl_mem in, int iw, int ih, cl_command_queue q);                      /* oclrect.cl:373 */
void rd_rect_reduceLS(cl_mem out, cl_mem boundary, cl_mem lsid, int iw, int ih, int nentry, cl_command_queue q); /* oclrect.cl:427 */

/* L3 introspection: device buffers of an oclrect_t by the reference's names ("buf0".."buf5", "tmp0".."tmp5",
 * "iobuf0", "iobuf1", "ioBig0", "ioBig1"; oclrect.c:51-53), and the host tail on caller-provided arrays. */
cl_mem rd_oclrect_buffer(struct oclrect_t *thiz, const char *name);
/* run genGPUTask's device schedule only (no read-back, no host tail); stop_step > 0: the operator-level replay stopped at
 * that step of SURVEY.md 10.1; 0: the production schedule; < 0: the production schedule stopped after stage -stop_step */
void   rd_oclrect_run_device(struct oclrect_t *thiz, const uint8_t *imgData, int ws, int stop_step);
/* executeCPUTask (oclrect.c:1049) on host arrays: ls = list incl. header, segid = plane, votes = int[nentry][5].  Pure host code,
 * kept as the checker of the device tail (the pipeline itself runs executeCPUTask on the device, rd_gtail.cu) */
rect_t *rd_rect_tail(const linesegment_t *ls, const int32_t *segid, const int32_t *votes, int iw, int ih, double tanAOV);
/* executeCPUTask (oclrect.c:1049) as a device operator on caller-owned buffers (the three read-backs of oclrect.c:371-376 as
 * cl_mems): segment list incl. header, region map, vote table.  Synchronous; returns a malloc()ed list, element 0 = header */
rect_t *rd_rect_tail_device(cl_mem lsList, cl_mem segid, cl_mem votes, int iw, int ih, double tanAOV, cl_command_queue q);
void   rd_free(void *p);

/* Frame-batch engine: `nctx` pipeline objects on one device, each with its own stream, buffers for
 * `frames_per_launch` frames and a host thread.  Every kernel launch processes up to frames_per_launch independent
 * frames.  rd_batch_run detects rectangles in `nframes` host frames (BGR8, stride ws, frame i at frames +
 * i*frame_stride) and returns one malloc()ed rect_t list per frame in out[i] (same format as oclrect_executeOnce).
 * Each frame is processed as by a freshly created oclrect_t (no state carried between frames). */
typedef struct rd_batch rd_batch;
rd_batch *rd_batch_create(int device, int iw, int ih, int nctx, int frames_per_launch);
void rd_batch_destroy(rd_batch *b);
void rd_batch_run(rd_batch *b, const uint8_t *frames, size_t frame_stride, int ws, int nframes, double tanAOV, rect_t **out);
/* NV12 input: decoded-video frames - Y plane with a row stride of ystride bytes, the interleaved half-resolution UV
 * plane behind it - in host or device memory (detected).  The YUV -> BGR step of the reference's video front end
 * (cv::VideoCapture, vidrect.cpp:160-166; integer BT.601 as OpenCV's COLOR_YUV2BGR_NV12) is fused into the first kernel. */
void rd_batch_run_nv12(rd_batch *b, const void *frames, size_t frame_stride, int ystride, int nframes, double tanAOV, rect_t **out);
rect_t *rd_oclrect_executeOnceNV12(struct oclrect_t *thiz, const uint8_t *nv12, int ystride, double tanAOV);
/* device-resident variant: frames already in device memory (read in place).  out == NULL runs everything on the device but
 * builds no host lists. */
void rd_batch_run_device(rd_batch *b, const void *dframes, size_t frame_stride, int ws, int nframes, double tanAOV, rect_t **out);
/* n lists as returned above -> one malloc()ed array holding all their entries (headers dropped, list order), counts[i] = entries of
 * list i; the lists are freed */
rect_t *rd_rect_lists_flatten(rect_t **lists, int n, int32_t *counts);

/* labelMergeMain (oclrect.cl:300-334) depends on the order of its work-items.  Default (0): the schedule-independent fixed point of
 * its adopt rule.  1: the reference's FIRST pass is replayed exactly in raster order (a latency-bound row wavefront, ~3 ms per
 * 1280x720 frame) and the fixed point is taken from there - the region map of the reference's sequential run (oracle/_ref) bit
 * for bit on all but one frame of the sweeps.  Process-wide; initial value from the environment variable RD_MERGE_REPLAY; read when
 * a task is enqueued (a captured CUDA graph of another mode is re-captured). */
void rd_set_merge_replay(int on);
int rd_get_merge_replay(void);
/* host-side accounting of the last rd_batch_run, summed over the pipeline objects' driver threads: out_ms[0] = time spent
 * waiting for the device, out_ms[4] = wall time spent copying the read-back records into rect_t lists; [1..3] are reserved (0) */
void rd_batch_stage_ms(rd_batch *b, double out_ms[5]);

#ifdef __cplusplus
}
#endif
#endif
