/* CL/cl.h - type-compatibility header for rectdetect_b200.
 *
 * NOT an OpenCL implementation.  The reference's public headers (oclimgutil.h, oclpolyline.h, oclrect.h,
 * oclhelper.h) and its demo programs (rect.cpp, poly.cpp, vidrect.cpp) spell their arguments with OpenCL
 * handle types.  This header gives those names a CUDA-backed meaning so that the same call sites compile
 * and link against librectdetect_b200.so with no OpenCL runtime present:
 *
 *   cl_device_id      -> a CUDA device ordinal
 *   cl_context        -> the primary context of that device
 *   cl_command_queue  -> one in-order cudaStream_t
 *   cl_mem            -> {device pointer, size}
 *   cl_event          -> a reference-counted cudaEvent_t
 *   cl_kernel/program -> unused opaque pointers (kernels are compiled ahead of time for sm_100a)
 *
 * Only the handful of entry points the reference's programs call directly are provided (rect.cpp:60-64,
 * poly.cpp:87-131, vidrect.cpp:120-127).
 */
#ifndef RECTDETECT_B200_CL_COMPAT_H
#define RECTDETECT_B200_CL_COMPAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int8_t   cl_char;
typedef uint8_t  cl_uchar;
typedef int16_t  cl_short;
typedef uint16_t cl_ushort;
typedef int32_t  cl_int;
typedef uint32_t cl_uint;
typedef int64_t  cl_long;
typedef uint64_t cl_ulong;
typedef float    cl_float;
typedef double   cl_double;
typedef cl_uint  cl_bool;
typedef cl_ulong cl_bitfield;
typedef cl_bitfield cl_mem_flags;
typedef cl_bitfield cl_command_queue_properties;

typedef struct rd_cl_platform *cl_platform_id;
typedef struct rd_cl_device   *cl_device_id;
typedef struct rd_cl_context  *cl_context;
typedef struct rd_cl_queue    *cl_command_queue;
typedef struct rd_cl_mem      *cl_mem;
typedef struct rd_cl_event    *cl_event;
typedef struct rd_cl_kernel   *cl_kernel;
typedef struct rd_cl_program  *cl_program;

#define CL_SUCCESS                      0
#define CL_INVALID_VALUE              (-30)
#define CL_INVALID_MEM_OBJECT         (-38)
#define CL_INVALID_COMMAND_QUEUE      (-36)
#define CL_MEM_OBJECT_ALLOCATION_FAILURE (-4)
#define CL_OUT_OF_RESOURCES           (-5)
#define CL_FALSE 0
#define CL_TRUE  1

#define CL_MEM_READ_WRITE      (1 << 0)
#define CL_MEM_WRITE_ONLY      (1 << 1)
#define CL_MEM_READ_ONLY       (1 << 2)
#define CL_MEM_USE_HOST_PTR    (1 << 3)
#define CL_MEM_ALLOC_HOST_PTR  (1 << 4)
#define CL_MEM_COPY_HOST_PTR   (1 << 5)
#define CL_MEM_HOST_NO_ACCESS  (1 << 9)
#define CL_QUEUE_PROFILING_ENABLE (1 << 1)

/* rect.cpp:64 / poly.cpp:47 / vidrect.cpp:127 */
cl_command_queue clCreateCommandQueue(cl_context context, cl_device_id device, cl_command_queue_properties properties, cl_int *errcode_ret);
cl_int clReleaseCommandQueue(cl_command_queue queue);
cl_int clReleaseContext(cl_context context);
cl_int clFlush(cl_command_queue queue);
cl_int clFinish(cl_command_queue queue);
/* poly.cpp:92-103 : flags are CL_MEM_READ_WRITE [| CL_MEM_COPY_HOST_PTR] or CL_MEM_HOST_NO_ACCESS */
cl_mem clCreateBuffer(cl_context context, cl_mem_flags flags, size_t size, void *host_ptr, cl_int *errcode_ret);
cl_int clReleaseMemObject(cl_mem mem);
cl_int clReleaseEvent(cl_event ev);
cl_int clRetainEvent(cl_event ev);
/* poly.cpp:127-129 ; blocking_read CL_TRUE synchronises the queue */
cl_int clEnqueueReadBuffer(cl_command_queue queue, cl_mem buffer, cl_bool blocking_read, size_t offset, size_t size, void *ptr,
                           cl_uint num_events_in_wait_list, const cl_event *event_wait_list, cl_event *event);
cl_int clEnqueueWriteBuffer(cl_command_queue queue, cl_mem buffer, cl_bool blocking_write, size_t offset, size_t size, const void *ptr,
                            cl_uint num_events_in_wait_list, const cl_event *event_wait_list, cl_event *event);

#ifdef __cplusplus
}
#endif
#endif
