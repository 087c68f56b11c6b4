// the reference PLY reader with quant::set_bounds (formats/ply/reader.cc:428) routed to the GPU
#include <cstdint>
#include <cstddef>
#include "structs/mesh.h"
#include "structs/quant.h"
#include "bridge.h"
#define set_bounds set_bounds_b200
#include "formats/ply/reader.cc"
