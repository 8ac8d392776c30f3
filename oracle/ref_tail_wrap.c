/* ref_tail_wrap.c - runs the REFERENCE's own host tail (executeCPUTask, oclrect.c:1049-1226, with its helpers
 * oclrect.c:385-1045, vec234.h, egbuf.h and helper.c's ArrayMap) on caller-provided arrays.
 *
 * Test infrastructure (oracle/): this file is compiled TOGETHER with the reference's oclrect.c - included below from
 * where it lies under /root/reference (include path set by oracle/Makefile; nothing of the reference is copied into
 * this repository) - against include/CL/cl.h of this repository (a type-compatibility header) and linked with the
 * reference's helper.c.  Output: oracle/_ref/librd_ref_tail.so.  It pins the oracle's restatement of the tail
 * (ora_tail.cpp) and the product's tail (rd_tail.cpp) to the reference's code: tests/test_ref_tail.py.
 *
 * oclrect.c also contains the device-side schedule (init_oclrect, genGPUTask), which needs an OpenCL runtime; those
 * functions are never called here, the symbols they reference are satisfied by the aborting stubs below. */
#include "oclrect.c"

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
