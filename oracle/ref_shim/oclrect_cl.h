/* Stand-in for the header the reference's build GENERATES from oclrect.cl (cltoheader, CMakeLists.txt:68-70): the
 * OpenCL program text.  Here the "source" is the tag by which the host runtime of oracle/_ref/librd_ref.so
 * (ref_cl_rt.cpp) finds the kernels translated from oclrect.cl; the reference-tail build never looks at it. */
static const char *source = "rect";
