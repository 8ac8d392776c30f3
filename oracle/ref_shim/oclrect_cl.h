/* Stand-in for the header the reference's build GENERATES from oclrect.cl (cltoheader, CMakeLists.txt:68-70): the
 * OpenCL program text.  The host tail never looks at it, so the reference-tail build (oracle/Makefile, target _ref)
 * gives oclrect.c an empty program. */
static const char *source = "";
