// bridge.h -- host side of the drop-in: marshals the reference's mesh::Mesh into the C ABI of
// include/harry_b200.h and replays / drains the symbol streams through the reference's own
// io::writer / io::reader.  Compiled against the reference headers where they lie (never copied).
#pragma once

#include <cstdint>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

#include "structs/mesh.h"
#include "structs/quant.h"

#include "../../include/harry_b200.h"

namespace b200 {

// one process-wide device context (GPU chosen with HARRY_B200_DEVICE, default 0)
hb_ctx *context();
[[noreturn]] void fail(const char *what);

struct FlatMesh {
	hb_mesh_desc desc;
	std::vector<hb_list_desc> lists;
	std::vector<int32_t> off_face, off_corner, off_vtx;
};

hb_list_desc describe_list(mesh::attr::Attr &attr);
// pointers alias the mesh's own vectors (structs/attr.h:103-109, structs/conn.h:80)
void flatten(mesh::Mesh &mesh, const std::vector<mesh::conn::fepair> &order, const std::vector<mesh::conn::fepair> *order_f, FlatMesh &out);

}

namespace quant {
// same signatures as quant::set_bounds (structs/quant.h:39-44) / quant::requant (:222-242)
void set_bounds_b200(mesh::attr::Attrs &attrs);
// what a reader does once all faces are in: twin matching of the half-edges on the GPU (replaces the per-corner
// hash join of mesh::Builder, structs/conn.h:164-214, which the reader TUs switch off with Builder::noautomerge,
// structs/mesh.h:49-52), then set_bounds
void finish_read_b200(mesh::Mesh &mesh);
void requant_b200(mesh::attr::Attrs &attrs, const std::vector<Quant> &quant, bool clear);
}
