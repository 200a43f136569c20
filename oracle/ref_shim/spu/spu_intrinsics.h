/* spu_intrinsics.h — stand-in for the Cell SDK header of that name (absent here).  TEST INFRASTRUCTURE ONLY.
 * cell/spu/trace_spu.c_ (C) uses one channel write and the MFC calls declared in spu_mfcio.h; nothing SIMD.
 * cell/spu/trace_spu.cpp (C++) also uses the `vector` type keyword and three SIMD intrinsics, stated here with
 * GCC's generic vector extension (four IEEE binary32 lanes, element-wise, never contracted under -ffp-contract=off):
 *   spu_splats(f) ....... all four lanes = f
 *   spu_cmpgt(a, b) ..... per lane: all ones if a > b, else 0
 *   spu_sel(a, b, m) .... per bit: m ? b : a                                  (SPU ISA: selb) */
#ifndef YV_REF_SHIM_SPU_INTRINSICS_H
#define YV_REF_SHIM_SPU_INTRINSICS_H
#define MFC_WrTagMask 22
#define spu_writech(channel, value) ((void)(channel), (void)(value))

#ifdef __cplusplus
#define vector __attribute__((vector_size(16)))        /* `vector float`, `vector unsigned int` */
typedef float yv_shim_f4 __attribute__((vector_size(16)));
typedef unsigned int yv_shim_u4 __attribute__((vector_size(16)));
static inline yv_shim_f4 spu_splats(float f) { yv_shim_f4 v = { f, f, f, f }; return v; }
static inline yv_shim_u4 spu_cmpgt(yv_shim_f4 a, yv_shim_f4 b) { return (yv_shim_u4)(a > b); }
static inline yv_shim_f4 spu_sel(yv_shim_f4 a, yv_shim_f4 b, yv_shim_u4 m) {
  return (yv_shim_f4)((((yv_shim_u4)a) & ~m) | (((yv_shim_u4)b) & m));
}
#endif
#endif
