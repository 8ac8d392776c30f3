/* rd_oracle.h - CPU oracle for the rectdetect hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a CPU restatement of the reference's OpenCL kernels (oclimgutil.cl, oclrect.cl,
 * oclpolyline.cl), of its launch schedules (oclrect.c:235-381 genGPUTask, oclpolyline.c:218-309
 * oclpolyline_execute, oclimgutil.c:227-273, poly.cpp:104-123) and of its host tail
 * (oclrect.c:385-1226 executeCPUTask).  It exists to check the CUDA path; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product library never links, imports or calls anything in this directory.
 *
 * PARITY STATUS: PINNED TO THE REFERENCE RUNNING HERE.  One kernel (labelMergeMain) depends on the order of the reference's own
 * work-items: by default the oracle computes a schedule-independent fixed point of its adopt rule, with ora_set_merge_replay(1) it
 * replays the reference's first pass in raster order first (bit-exact to the reference's kernel) and then equals the reference's
 * sequential run on all rectangles of the sweeps.  The reference ships no tests, golden vectors or fixtures and no
 * OpenCL runtime exists in this image, but the reference itself does run: `make _ref` compiles its host code
 * (helper.c, oclhelper.c, oclimgutil.c, oclpolyline.c, oclrect.c - unmodified, from /root/reference) together with its
 * three OpenCL C kernel files compiled as C++ (cl_translate.py rewrites only the vector-literal syntax, cl_compat.h supplies
 * the OpenCL C built-ins) over a synchronous host runtime (ref_cl_rt.cpp) into oracle/_ref/librd_ref.so; every NDRange runs
 * its work-items in raster order on one thread - one legal schedule of the reference.  Against it (tests/test_ref_device.py,
 * tests/golden/ref_device_golden.json):
 *   - Stage A and Stage B up to the merge mask (genGPUTask steps 1-16): every plane bit-exact, floats included;
 *   - Stage C (oclpolyline_execute, all 116 launches): every plane, the segment-id map and the LS_t list bit-exact;
 *   - calcSize, markBoundary, label8x, reduceLS (vote table) on identical inputs: bit-exact;
 *   - executeCPUTask (host tail): bit-exact (tests/test_ref_tail.py, tests/golden/ref_tail_golden.json);
 *   - despeckle2 (Q3) updates its labels in place: the oracle evaluates it in raster order, i.e. exactly as the reference run
 *     does - bit-exact on identical inputs;
 *   - labelMergeMain (Q6') is ORDER DEPENDENT in the reference (directed adopt rule gated on the current labels).  The
 *     oracle fixes a deterministic fixed point of the same rule (two-directional pairs united, one-directional pairs united
 *     where the source's component label is smaller; ora_rect.cpp) whose distance from the sequential schedule is tested (a
 *     handful to a few hundred interior pixels per frame); with that one kernel swapped for the reference's the oracle
 *     reproduces the reference's region map bit-exactly.  ora_set_merge_replay(1): labelxPreprocess + the FIRST labelMergeMain pass
 *     as the reference's kernel runs them in raster order (whole label plane bit-exact), then the same fixed point seeded with
 *     that plane: label plane identical to the reference's 8 passes on 24 of 27 sweep frames (3 / 1 / 87 px on the others), region map on 31 of 33, all 249
 *     rectangles identical (profiles/r04t_*).
 * Every other place where the reference is schedule-dependent (atomic arrival order, in-place races, vote-slot claims)
 * is resolved the way the raster-order schedule resolves it; each is marked "CANONICAL" in the sources and listed in
 * DESIGN.md section "Canonical semantics".  Vendor-defined OpenCL built-ins (rsqrt, hypot, distance, FP contraction) follow
 * the choices stated there (Q14-Q18) in the oracle, in cl_compat.h and in the CUDA kernels alike.
 */
#ifndef RD_ORACLE_H
#define RD_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* == linesegment_t of the reference (oclpolyline.h:74-83) == */
typedef struct ora_ls_t {
  float x0, y0, x1, y1;
  int32_t startIndex, endIndex, leftPtr, rightPtr, startCount, endCount, maxDist, polyid, npix, level;
} ora_ls_t;

/* == rect_t of the reference (oclrect.h:5-15): 176 bytes, element 0 is a header == */
typedef struct ora_rect_t {
  union {
    struct { double c2[4][2]; double c3[4][3]; double value; uint32_t status; };
    int nItems;
  };
} ora_rect_t;

typedef struct ora_stats_t {
  int label8x_seq_passes;      /* passes the reference's label8xMain needs under a sequential schedule (max over calls) */
  int label8x_calls;
  int labelpl_components;
  int mkpl_ties;               /* exact arg-max ties seen in mkpl */
  int mkpl_iterations_live;    /* how many of the 15 split iterations did work */
  int vote_collisions;         /* reduceLS slots wanted by more than one lsid */
  int vote_slots;              /* occupied slots */
  int n_ls;                    /* lsList[0] count after polyline */
  int ls_overflow;             /* "Too many linesegments" events */
} ora_stats_t;

void ora_set_threads(int n);
int  ora_get_threads(void);
void ora_get_stats(ora_stats_t *out);
void ora_reset_stats(void);

/* ---- Stage A operators: one function per oclimgutil_* wrapper (oclimgutil.c:140-319) ---- */
void ora_clear(int32_t *out, int size_bytes);
void ora_copy(int32_t *out, const int32_t *in, int size_bytes);
void ora_cast_i_f(int32_t *out, const float *in, float scale, int size);
void ora_cast_c_i(int8_t *out, const int32_t *in, int size);
void ora_threshold_i_i(int32_t *out, const int32_t *in, int vlow, int threshold, int vhigh, int size);
void ora_threshold_f_f(float *out, const float *in, float vlow, float threshold, float vhigh, int size);
void ora_convert_plab_bgr(uint32_t *out, const uint8_t *in, int iw, int ih, int ws);   /* runs bgr2plab (Q9) */
void ora_unpack_f_f_f_plab(float *o0, float *o1, float *o2, const uint32_t *in, int iw, int ih);
void ora_pack_plab_f_f_f(uint32_t *out, const float *i0, const float *i1, const float *i2, int iw, int ih);
void ora_iirblur_f_f(float *obuf, const float *ibuf, float *tmp0, float *tmp1, int r, int iw, int ih);
void ora_edgevec_f2_f(float *out_xy, const float *in, int iw, int ih);
/* NV12 -> BGR8 as OpenCV's COLOR_YUV2BGR_NV12 (the video front end of vidrect.cpp:160-166); ys = row stride of the Y / UV planes */
void ora_nv12_to_bgr(uint8_t *bgr, const uint8_t *nv12, int iw, int ih, int ws, int ys);
/* operators no configured path enqueues (oclimgutil.h:86-94): visualisers, alternative edge / thinning kernels */
void ora_edgevec_f2_plab(float *out_xy, const uint32_t *in, int iw, int ih);                      /* oclimgutil.cl:354 */
void ora_edge_f_f(float *out, const float *in, int iw, int ih);                                   /* oclimgutil.cl:439 */
void ora_thincubic_f_f_f2(float *out, const float *in, const float *vxy, int iw, int ih);         /* oclimgutil.cl:473 */
void ora_convert_bgr_plab(uint8_t *out, const uint32_t *in, int iw, int ih, int ws);              /* oclimgutil.cl:264 plab2bgr (Q9) */
void ora_convert_bgr_lumaf(uint8_t *out, const float *in, float f, int iw, int ih, int ws);       /* oclimgutil.cl:283 */
void ora_convert_bgr_labeli(uint8_t *out, const int32_t *in, int bgc, int iw, int ih, int ws);    /* oclimgutil.cl:291 */
void ora_edge_f_plab(float *out, const uint32_t *in, int iw, int ih);
void ora_thinthres_f_f_f2(float *out, const float *in, const float *vxy, int iw, int ih);
/* CANONICAL: converged labels.  Returns number of sequential passes the reference kernel needed. */
int  ora_label8x_int_int(int32_t *out, const int32_t *in, int32_t *tmp, int bgc, int iw, int ih);
void ora_calcStrength(int32_t *out, const float *edge, const int32_t *label, int iw, int ih);
void ora_filterStrength(int32_t *labelinout, const int32_t *str, int thre, int iw, int ih);

/* ---- Stage B kernels of oclrect.cl, exposed one by one for operator-level parity ---- */
void ora_rect_simpleJunction(int32_t *out, const int32_t *in, int iw, int ih);
void ora_rect_simpleConnect(int32_t *out, const int32_t *in, int iw, int ih);
void ora_rect_stringify(int32_t *out, const int32_t *in, int mod2, int iw, int ih);
void ora_rect_blblur0(uint32_t *out, const int8_t *edge, const uint32_t *in, int iw, int ih);
void ora_rect_blblur1(uint32_t *out, const int8_t *edge, const uint32_t *in, int iw, int ih);
void ora_rect_quantize(uint32_t *out, const uint32_t *in, int n0, int n1, int n2, int iw, int ih);
void ora_rect_despeckle(uint32_t *out, const uint32_t *in, const float *edge, int iw, int ih);
void ora_rect_mkMergeMask0(int32_t *out, const int32_t *junction, int iw, int ih);
void ora_rect_mkMergeMask1(int32_t *inout, const int32_t *junction, int iw, int ih);
/* labelxPreprocess + labelMergeMain x8 -> CANONICAL converged symmetric merge */
/* which form of the merge labelling the oracle computes (process-wide): 0 = the schedule-independent fixed point (default, = the CUDA
 * path by default), 1 = the reference's first pass replayed in raster order, then the fixed point (= the CUDA path with RD_MERGE_REPLAY=1) */
void ora_set_merge_replay(int on);
/* the fixed point seeded with the label plane found in `label` (a state of the reference's plane after one of its passes) */
void ora_rect_labelMerge_seeded(int32_t *label, const int32_t *pix, const int32_t *mask, const int32_t *edge, int iw, int ih);
int ora_get_merge_replay(void);
/* labelxPreprocess + the first labelMergeMain pass in raster order (= the reference kernel run sequentially once) */
void ora_rect_labelMerge_first_pass(int32_t *label, const int32_t *pix, const int32_t *mask, const int32_t *edge, int iw, int ih);
void ora_rect_labelMerge(int32_t *label, const int32_t *pix, const int32_t *mask, const int32_t *edge, int iw, int ih);
void ora_rect_calcSize(int32_t *out, const int32_t *label, int iw, int ih);
/* in place, work-items in raster order (CANONICAL Q3 = the reference run) */
void ora_rect_despeckle2(int32_t *labelinout, const int32_t *size, int thre, int iw, int ih);
void ora_rect_markBoundary(int32_t *out, const int32_t *in, int iw, int ih);
void ora_rect_reduceLS(int32_t *out, const int32_t *boundary, const int32_t *lsid, int iw, int ih, int nentry);

/* ---- Stage C: oclpolyline_execute (oclpolyline.c:218), same argument order minus the CL handles.
 *      stop_step: 0 = run everything, k>0 = return after step k of SURVEY.md 10.2 (for intermediates). ---- */
void ora_polyline_execute(ora_ls_t *lsList, int lsListSize, int32_t *lsIdOut, const int32_t *in, int32_t *tmpBig,
                          int32_t *tmp0, int32_t *tmp1, int32_t *tmp2, int32_t *tmp3, int32_t *tmp4, int32_t *tmp5,
                          float minerror, int sizeThre, int iw, int ih, int stop_step);

/* ---- L3: the oclrect_t object (oclrect.c:41-135) with the reference's buffer aliasing ---- */
typedef struct ora_rect ora_rect;
ora_rect *ora_rect_create(int iw, int ih);
void      ora_rect_destroy(ora_rect *o);
/* genGPUTask (oclrect.c:235-381).  stop_step: 0 = all, k>0 = return after step k of SURVEY.md 10.1 */
void      ora_rect_gpu_task(ora_rect *o, const uint8_t *img, int ws, int stop_step);
/* names: "buf0".."buf5", "tmp0".."tmp5", "iobuf0", "iobuf1", "ioBig0", "ioBig1" */
void     *ora_rect_buffer(ora_rect *o, const char *name);
/* executeCPUTask (oclrect.c:1049-1226) on the object's buffers; result is malloc()ed, element 0 = header */
ora_rect_t *ora_rect_cpu_task(ora_rect *o, double tanAOV);
/* oclrect_executeOnce (oclrect.c:1230) */
ora_rect_t *ora_rect_execute_once(ora_rect *o, const uint8_t *img, int ws, double tanAOV);
/* the host tail on caller-provided arrays (full-size vote table as the reference reads it back) */
ora_rect_t *ora_tail(const ora_ls_t *ls, const int32_t *segid, const int32_t *votes, int iw, int ih, double tanAOV);
void ora_free(void *p);
/* per-stage wall-clock of the last ora_rect_gpu_task / cpu_task, seconds: A, B, C, D, tail */
void ora_rect_last_times(ora_rect *o, double out[5]);

/* ---- poly.cpp:104-123 pipeline (config 1): fills lsId map (mem0) and LS list ---- */
void ora_poly_frame(const uint8_t *img, int ws, int iw, int ih, float minerror, int sizeThre, int strengthThre,
                    int32_t *lsIdOut, ora_ls_t *lsListOut /* iw*ih*16 bytes */, float *thinOut /* may be NULL */);

/* ---- small pure functions exported for known-answer tests ---- */
uint32_t ora_srgb2plab(int b, int g, int r);
uint32_t ora_packlab(float l, float a, float b);
void     ora_unpacklab(uint32_t plab, float out[3]);
int      ora_mirror1(int x, int iw);
int      ora_repeat1(int x, int iw);
uint64_t ora_xrandom(uint64_t s);
int32_t  ora_rand_at(int x, uint64_t seed);
void     ora_clip_line(double x0, double y0, double x1, double y1, double xmin, double ymin, double xmax, double ymax, double out[4]);
void     ora_intersection2(const double u[4], const double v[4], double out[2]);
/* poseEstimation (oclrect.c:590) on 4 corner points given in image order */
void     ora_pose(const double corners[4][2], int iw, int ih, double tanAOV, ora_rect_t *out);

#ifdef __cplusplus
}
#endif
#endif
