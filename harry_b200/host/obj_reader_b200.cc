// the reference OBJ reader with quant::set_bounds (formats/obj/reader.cc:1253) routed to the GPU
#include <cstdint>
#include <cstddef>
#include "structs/mesh.h"
#include "structs/quant.h"
#include "bridge.h"
#define set_bounds set_bounds_b200
#include "formats/obj/reader.cc"
