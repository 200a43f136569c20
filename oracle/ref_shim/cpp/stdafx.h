/* stdafx.h — stand-in for the reference's precompiled header (cpp/stdafx.h, absent from the snapshot).
 * TEST INFRASTRUCTURE ONLY: lets cell/ppu_renderer.cpp and cell/spu/trace_spu.cpp compile unmodified, from where
 * they lie, into oracle/_ref (oracle/Makefile, target ref). It brings in what those files use without including it
 * themselves: the C/C++ library, the reference's own vector types (nest/include/geometry/primitives/point.h — the
 * real header, found through -I $(REFERENCE)/nest/include), `uint`, shared_ptr, min/max. */
#ifndef YV_REF_SHIM_STDAFX_H
#define YV_REF_SHIM_STDAFX_H

#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <math.h>
#include <stdio.h>
#include <assert.h>
#include <sys/time.h>
#include <algorithm>
#include <utility>
#include <iostream>
#include <fstream>
#include <memory>

typedef unsigned int uint;

#include <geometry/primitives/point.h>      /* the reference's cg::point_t (nest/include) */

using cg::point_3f;
using cg::point_2i;
using std::shared_ptr;
using std::min;
using std::max;

#define GLOBAL_FUNC                         /* __host__ __device__ on the reference's CUDA build */

#ifdef TARGET_SPU
#include <spu_intrinsics.h>
#include <spu_mfcio.h>
#endif

#endif
