/* libmisc.h — stand-in for the Cell SDK header of that name (absent).  TEST INFRASTRUCTURE ONLY.
 * cell/spu/trace_spu.cpp includes it and uses nothing from it. */
#ifndef YV_REF_SHIM_LIBMISC_H
#define YV_REF_SHIM_LIBMISC_H
#endif
