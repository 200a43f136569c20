/* data.h — stand-in for the reference's cell/data.h, which is not in the snapshot.  TEST INFRASTRUCTURE ONLY.
 *
 * cell/spu/trace_spu.c_ (the reference's first, scalar, double-precision SPU tracer) includes "../data.h" for its node
 * type and its per-SPE job record. The fields below are exactly the ones the two files that use them touch:
 *   Node.type[8], Node.children[8], LEAF / BRANCHING / EMPTY ... cell/trace.c_:25-45, cell/spu/trace_spu.c_:42-52
 *   spu_context.root/width/heigth/dx/dy/x/y/result .............. cell/trace.c_:86-96, cell/spu/trace_spu.c_:121-133
 * children[] is an int there (32-bit PPU pointers are stored in it, cell/trace.c_:32); tests place the pool below 2 GiB
 * so the same holds here. Used only to compile that file, unmodified, into oracle/_ref (oracle/Makefile, target ref).
 */
#ifndef YV_REF_SHIM_DATA_H
#define YV_REF_SHIM_DATA_H

enum { EMPTY = 0, LEAF = 1, BRANCHING = 2 };

typedef struct Node {
  int type[8];
  int children[8];
} Node;

typedef struct spu_context {
  Node *root;
  int width, heigth;
  int dx, dy, x, y;
  int *result;
} spu_context;

#endif
