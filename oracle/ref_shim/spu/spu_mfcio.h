/* spu_mfcio.h — stand-in for the Cell SDK's MFC DMA interface (absent here).  TEST INFRASTRUCTURE ONLY.
 * A blocking DMA "get" from main memory into the SPU's local store (cell/spu/trace_spu.c_:16-17,118-119) is a memcpy
 * on a machine with one address space; a "put" is the memcpy the other way. Effective addresses are 32-bit there. */
#ifndef YV_REF_SHIM_SPU_MFCIO_H
#define YV_REF_SHIM_SPU_MFCIO_H
#include <stdint.h>
#include <string.h>
#define MFC_GET_CMD 0x40
#define MFC_PUT_CMD 0x20
#define MFC_TAG_UPDATE_ALL 2
static inline void spu_mfcdma32(volatile void *ls, unsigned int ea, unsigned int size, unsigned int tag, unsigned int cmd) {
  (void)tag;
  if (cmd == MFC_GET_CMD) memcpy((void *)ls, (const void *)(uintptr_t)ea, size);
  else memcpy((void *)(uintptr_t)ea, (const void *)ls, size);
}
static inline unsigned int spu_mfcstat(unsigned int type) { (void)type; return 0; }
static inline unsigned int mfc_tag_reserve(void) { return 0; }
#endif
