// main_b200.cc -- the reference CLI (main.cc, compiled from the reference tree unchanged) with
// quant::requant (main.cc:108) routed to the GPU.  Same flags, same output.
#include <cstdint>
#include <cstddef>
#include "structs/mesh.h"
#include "structs/quant.h"
#include "bridge.h"
#define requant requant_b200
#include "main.cc"
