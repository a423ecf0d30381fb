/* TEST INFRASTRUCTURE — type-only stand-in for <CL/cl_platform.h>.
 *
 * The reference's vector type is `typedef cl_float4 Vector3` (vector3_cl.h:14) and
 * vector3_cl.h:8-12 includes <CL/cl_platform.h> purely for the type names; no OpenCL
 * runtime is touched on the native photon-mapping path.  This container has no OpenCL
 * headers, so the oracle build (oracle/Makefile) puts this directory on the include
 * path when it compiles the unmodified reference sources out of /root/reference.
 *
 * The real header also drags in <stddef.h>/<stdint.h> and, under SSE, the intrinsics
 * headers, which is how the reference gets size_t (helpers.h:6) and rand()/RAND_MAX
 * (vector3_cl.c:107, via xmmintrin.h -> mm_malloc.h -> stdlib.h).  Mirror that.
 */
#ifndef FMGI_CL_PLATFORM_STUB_H
#define FMGI_CL_PLATFORM_STUB_H

#include <stddef.h>
#include <stdint.h>
#if defined(__SSE__)
#include <xmmintrin.h>
#endif
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

typedef int32_t  cl_int;
typedef uint32_t cl_uint;
typedef float    cl_float;

typedef union {
    cl_float s[4] __attribute__((aligned(16)));
    struct { cl_float x, y, z, w; };
} cl_float4;
typedef cl_float4 cl_float3;

typedef union {
    cl_int s[4] __attribute__((aligned(16)));
    struct { cl_int x, y, z, w; };
} cl_int4;
typedef cl_int4 cl_int3;

#endif
