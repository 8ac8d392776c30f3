/* ref_full_wrap.c - test-side access to the REFERENCE's rectangle detector running here in full.
 *
 * TEST INFRASTRUCTURE (oracle/).  Compiled TOGETHER with the reference's oclrect.c (included below from where it lies
 * under /root/reference; nothing of the reference is copied into this repository) and linked with the reference's
 * helper.c, oclhelper.c, oclimgutil.c, oclpolyline.c (unmodified), the reference's three .cl kernel files compiled as
 * C++ (cl_translate.py, cl_compat.h) and the host runtime ref_cl_rt.cpp into oracle/_ref/librd_ref.so.
 * The public entry points (init_oclrect, oclrect_executeOnce, oclimgutil_*, oclpolyline_execute ...) are the
 * reference's own; the functions below only expose what is private to oclrect.c: the device planes of an oclrect_t,
 * the two halves of executeOnce (genGPUTask, executeCPUTask) and the host tail on caller-provided arrays. */
#include "oclrect.c"

void *rd_ref_mem_ptr(cl_mem m);

/* host pointer of a device plane: "buf0".."buf5", "tmp0".."tmp5", "iobuf0/1", "ioBig0/1" (oclrect.c:53-55) */
void *rd_ref_rect_buffer(oclrect_t *thiz, const char *name) {
  assert(thiz->magic == MAGIC);
  if (!strncmp(name, "buf", 3) && name[3] >= '0' && name[3] < '0' + NBUF) return rd_ref_mem_ptr(thiz->buf[name[3] - '0']);
  if (!strncmp(name, "tmp", 3) && name[3] >= '0' && name[3] < '0' + NTMP) return rd_ref_mem_ptr(thiz->tmp[name[3] - '0']);
  if (!strncmp(name, "iobuf", 5) && name[5] >= '0' && name[5] < '2') return rd_ref_mem_ptr(thiz->iobuf[name[5] - '0']);
  if (!strncmp(name, "ioBig", 5) && name[5] >= '0' && name[5] < '2') return rd_ref_mem_ptr(thiz->ioBig[name[5] - '0']);
  return NULL;
}
/* back to the state init_oclrect leaves: every device plane zero (the reference keeps state across frames in buf[3], Q1).
 * The reference's helper layer hands out at most 1000 kernel ids per process (oclhelper.c KERNELIDMAX), 19 per
 * init_oclrect, so the tests reuse one object per frame size instead of creating fresh ones. */
size_t rd_ref_mem_bytes(cl_mem m);
void rd_ref_rect_reset(oclrect_t *thiz) {
  assert(thiz->magic == MAGIC);
  for (int i = 0; i < NBUF; i++) memset(rd_ref_mem_ptr(thiz->buf[i]), 0, rd_ref_mem_bytes(thiz->buf[i]));
  for (int i = 0; i < NTMP; i++) memset(rd_ref_mem_ptr(thiz->tmp[i]), 0, rd_ref_mem_bytes(thiz->tmp[i]));
  for (int i = 0; i < 2; i++) {
    memset(rd_ref_mem_ptr(thiz->iobuf[i]), 0, rd_ref_mem_bytes(thiz->iobuf[i]));
    memset(rd_ref_mem_ptr(thiz->ioBig[i]), 0, rd_ref_mem_bytes(thiz->ioBig[i]));
  }
}
/* the device half of oclrect_executeOnce (oclrect.c:1230-1246): genGPUTask on page 0, waited for */
void rd_ref_gen_gpu_task(oclrect_t *thiz, uint8_t *img, int ws) {
  cl_event ev = genGPUTask(thiz, img, 0, ws, thiz->queue, NULL);
  waitForEvent(ev);
  ce(clReleaseEvent(ev));
}
/* the host half on what genGPUTask read back into page 0 */
rect_t *rd_ref_cpu_task(oclrect_t *thiz, double tanAOV) { return executeCPUTask(thiz, 0, tanAOV); }

/* executeCPUTask only touches iw, ih and the three host arrays of the page (oclrect.c:1050-1126) */
rect_t *rd_ref_execute_cpu_task(const int32_t *lsList, const int32_t *votes, const int32_t *segid, int iw, int ih, double tanAOV) {
  oclrect_t t;
  memset(&t, 0, sizeof(t));
  t.magic = MAGIC;
  t.iw = iw;
  t.ih = ih;
  t.hostioBig[0][0] = (cl_int *)lsList;
  t.hostioBig[0][1] = (cl_int *)votes;
  t.hostiobuf[0][1] = (cl_int *)segid;
  return executeCPUTask(&t, 0, tanAOV);
}
void rd_ref_free(void *p) { free(p); }
