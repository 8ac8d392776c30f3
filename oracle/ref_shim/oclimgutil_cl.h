/* stands in for the cltoheader output: the "source" is the tag the host runtime (ref_cl_rt.cpp) finds the translated kernels by */
static const char *source = "imgutil";
