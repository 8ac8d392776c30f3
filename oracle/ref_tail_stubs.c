/* ref_tail_stubs.c - link-time stand-ins for the OpenCL runtime / device-side helper functions that the reference's
 * oclrect.c references from init_oclrect, genGPUTask and dispose_oclrect.  The reference-tail library
 * (oracle/_ref/librd_ref_tail.so) only ever calls executeCPUTask, so none of these may be reached. */
#include <stdio.h>
#include <stdlib.h>
#define STUB(name) void name(void) { fprintf(stderr, "rd_ref_tail: %s called - the reference-tail library has no device side\n", #name); abort(); }
STUB(allocatePinnedMemory) STUB(freePinnedMemory) STUB(ce) STUB(waitForEvent) STUB(getNextKernelID) STUB(runKernel2Dx)
STUB(simpleBuildProgram) STUB(simpleSetKernelArg)
STUB(clCreateBuffer) STUB(clCreateKernel) STUB(clCreateProgramWithSource) STUB(clEnqueueReadBuffer) STUB(clEnqueueWriteBuffer)
STUB(clFlush) STUB(clReleaseEvent) STUB(clReleaseKernel) STUB(clReleaseMemObject) STUB(clReleaseProgram)
STUB(oclimgutil_cast_c_i) STUB(oclimgutil_cast_i_f) STUB(oclimgutil_clear) STUB(oclimgutil_convert_plab_bgr) STUB(oclimgutil_edge_f_plab)
STUB(oclimgutil_edgevec_f2_f) STUB(oclimgutil_iirblur_f_f) STUB(oclimgutil_label8x_int_int) STUB(oclimgutil_pack_plab_f_f_f)
STUB(oclimgutil_thinthres_f_f_f2) STUB(oclimgutil_threshold_f_f) STUB(oclimgutil_threshold_i_i) STUB(oclimgutil_unpack_f_f_f_plab)
STUB(oclpolyline_execute)
