/* ref_shim/CL/cl.h - the OpenCL declarations the REFERENCE's host code (oclhelper.c, oclimgutil.c, oclpolyline.c,
 * oclrect.c) needs to compile unmodified for oracle/_ref/librd_ref.so.  TEST INFRASTRUCTURE (oracle/).
 * Types and the app-level entry points come from this repository's include/CL/cl.h; this header adds what only the
 * reference's own helper layer calls.  The functions are implemented by oracle/ref_cl_rt.cpp: a synchronous,
 * single-device host "runtime" whose kernels are the reference's .cl files compiled as C++ (cl_translate.py). */
#ifndef RD_REF_SHIM_CL_H
#define RD_REF_SHIM_CL_H
#include "../../../include/CL/cl.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef cl_bitfield cl_device_type;
typedef cl_bitfield cl_map_flags;
typedef cl_uint cl_device_info, cl_program_build_info, cl_kernel_info, cl_event_info, cl_profiling_info;
typedef intptr_t cl_context_properties;

#define CL_DEVICE_TYPE_ALL 0xFFFFFFFF
#define CL_DEVICE_NAME 0x102B
#define CL_DEVICE_VERSION 0x102F
#define CL_DEVICE_EXTENSIONS 0x1030
#define CL_PROGRAM_BUILD_LOG 0x1183
#define CL_KERNEL_FUNCTION_NAME 0x1190
#define CL_EVENT_COMMAND_EXECUTION_STATUS 0x11D3
#define CL_PROFILING_COMMAND_START 0x1282
#define CL_PROFILING_COMMAND_END 0x1283
#define CL_COMPLETE 0
#define CL_MAP_READ (1 << 0)
#define CL_MAP_WRITE (1 << 1)
#define CL_INVALID_WORK_GROUP_SIZE (-54)
#define CL_INVALID_KERNEL_NAME (-46)
#define CL_INVALID_ARG_INDEX (-49)
#define CL_INVALID_ARG_SIZE (-51)

cl_int clGetPlatformIDs(cl_uint n, cl_platform_id *platforms, cl_uint *nret);
cl_int clGetDeviceIDs(cl_platform_id p, cl_device_type t, cl_uint n, cl_device_id *devices, cl_uint *nret);
cl_int clGetDeviceInfo(cl_device_id d, cl_device_info what, size_t size, void *value, size_t *size_ret);
cl_context clCreateContext(const cl_context_properties *props, cl_uint n, const cl_device_id *devices,
                           void (*notify)(const char *, const void *, size_t, void *), void *user, cl_int *err);
cl_program clCreateProgramWithSource(cl_context c, cl_uint count, const char **strings, const size_t *lengths, cl_int *err);
cl_int clBuildProgram(cl_program p, cl_uint n, const cl_device_id *devices, const char *options, void (*notify)(cl_program, void *), void *user);
cl_int clGetProgramBuildInfo(cl_program p, cl_device_id d, cl_program_build_info what, size_t size, void *value, size_t *size_ret);
cl_int clReleaseProgram(cl_program p);
cl_kernel clCreateKernel(cl_program p, const char *name, cl_int *err);
cl_int clReleaseKernel(cl_kernel k);
cl_int clGetKernelInfo(cl_kernel k, cl_kernel_info what, size_t size, void *value, size_t *size_ret);
cl_int clSetKernelArg(cl_kernel k, cl_uint index, size_t size, const void *value);
cl_int clEnqueueNDRangeKernel(cl_command_queue q, cl_kernel k, cl_uint dim, const size_t *offset, const size_t *gws, const size_t *lws,
                              cl_uint nev, const cl_event *wait, cl_event *event);
void *clEnqueueMapBuffer(cl_command_queue q, cl_mem m, cl_bool blocking, cl_map_flags flags, size_t offset, size_t size,
                         cl_uint nev, const cl_event *wait, cl_event *event, cl_int *err);
cl_int clEnqueueUnmapMemObject(cl_command_queue q, cl_mem m, void *ptr, cl_uint nev, const cl_event *wait, cl_event *event);
cl_int clGetEventInfo(cl_event e, cl_event_info what, size_t size, void *value, size_t *size_ret);
cl_int clGetEventProfilingInfo(cl_event e, cl_profiling_info what, size_t size, void *value, size_t *size_ret);
#ifdef __cplusplus
}
#endif
#endif
