// the reference OBJ reader with the Builder's twin matching and quant::set_bounds (formats/obj/reader.cc:1253) routed to the GPU
#include <cstdint>
#include <cstddef>
#include "structs/mesh.h"
#include "structs/quant.h"
#include "bridge.h"
#define set_bounds(attrs) finish_read_b200(mesh)
// twin matching moves to the GPU (hb_twin_match inside finish_read_b200): the Builder's own per-corner hash join
// (structs/conn.h:201-214) is switched off right after the reader's single init_bindings call, before any face
#define init_bindings(a, b, c) init_bindings(a, b, c), builder.noautomerge()
#include "formats/obj/reader.cc"
