/* spu_intrinsics.h — stand-in for the Cell SDK header of that name (absent here).  TEST INFRASTRUCTURE ONLY.
 * cell/spu/trace_spu.c_ uses one channel write and the MFC calls declared in spu_mfcio.h; nothing SIMD. */
#ifndef YV_REF_SHIM_SPU_INTRINSICS_H
#define YV_REF_SHIM_SPU_INTRINSICS_H
#define MFC_WrTagMask 22
#define spu_writech(channel, value) ((void)(channel), (void)(value))
#endif
