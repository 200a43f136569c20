// harness_cell_app.cpp — the one symbol the reference's Cell application expects from its build system.
// TEST INFRASTRUCTURE ONLY. cell/spu_renderer.cpp:9 declares `extern spe_program_handle_t trace_spu;`: on the Cell the
// SPU program is embedded into the PPU executable under that name (cell/spu/Makefile:51 LIBRARY_embed). On the host the
// "program" is the function trace_spu_main linked into the same executable (oracle/ref_shim/spe/libspe2.h), and the
// handle is a placeholder. The stand-in shader's probe switch (oracle/ref_shim/cpp/shader.h) lives here too.
#include "stdafx.h"
#include <libspe2.h>
spe_program_handle_t trace_spu = { 0 };
extern "C" { int yv_ref_shader_probe = 0; }
