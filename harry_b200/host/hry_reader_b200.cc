// hry_reader_b200.cc -- the reference .hry reader (formats/hry/reader.cc, compiled unchanged) with
// AttrDecoder<io::reader>::decode (attrcode.h:534-550, call site reader.cc:192) specialized:
//   phase 1 (host, sequential): drain the arithmetic-coded symbol stream exactly in the order of
//           vtx_post / face_post / corner_post (attrcode.h:443-531) WITHOUT prediction -- regions,
//           type symbols, residual rows, history offsets -> binding tables.  Legal because the
//           adaptive models only adapt on symbols.
//   phase 2 (GPU): hb_attr_decode reconstructs all values from the residual rows.
#include <cstdint>
#include <cstddef>
#include <cstring>
#include "formats/hry/reader.h"
#include "formats/hry/common.h"
#include "formats/hry/attrcode.h"
#include "formats/hry/io.h"
#include "utils/progress.h"
#include "bridge.h"

namespace hry {
namespace attrcode {

template <>
template <>
void AttrDecoder<io::reader>::decode<progress::handle>(progress::handle &prog)
{
	mesh::Mesh &m = builder.mesh;
	const size_t nlists = m.attrs.size();
	std::vector<std::vector<uint8_t>> types(nlists);
	auto read_one = [&](mesh::listidx_t l, mesh::listidx_t a, mesh::vtxidx_t corner_vtx, bool corner) -> mesh::attridx_t {
		const AttrType t = rd.attr_type(l);
		types[l].push_back((uint8_t)t);
		mesh::attridx_t idx;
		switch (t) {
		case DATA:
			idx = cur_idx[l]++;
			rd.attr_data(m.attrs[l][idx], l); // the residual row stays in place; the GPU reconstructs it
			if (corner) lhist[a].insert(corner_vtx, idx);
			break;
		case HIST:
			idx = cur_idx[l] - 1 - rd.attr_ghist(l);
			if (corner) lhist[a].insert(corner_vtx, idx);
			break;
		default:
			idx = lhist[a].find(corner_vtx, rd.attr_lhist(l));
			break;
		}
		return idx;
	};

	prog.start(order.size());
	for (size_t i = 0; i < order.size(); ++i) { // vtx_post, attrcode.h:443-470
		const mesh::vtxidx_t v = m.conn.org(order[i]);
		const mesh::regidx_t r = rd.reg_vtx();
		builder.vtx_reg(v, r);
		for (mesh::listidx_t a = 0; a < m.attrs.num_bindings_vtx_reg(r); ++a)
			builder.bind_vtx_attr(v, a, read_one(m.attrs.binding_reg_vtxlist(r, a), a, 0, false));
		prog(i);
	}
	for (mesh::faceidx_t f = 0; f < m.attrs.num_face(); ++f) { // face_post :476-501, corner_post :502-531
		const mesh::regidx_t r = rd.reg_face();
		builder.face_reg(f, r);
		for (mesh::listidx_t a = 0; a < m.attrs.num_bindings_face_reg(r); ++a)
			builder.bind_face_attr(f, a, read_one(m.attrs.binding_reg_facelist(r, a), a, 0, false));
		for (int c = 0; c < m.conn.num_edges(f); ++c)
			for (mesh::listidx_t a = 0; a < m.attrs.num_bindings_corner_reg(r); ++a)
				builder.bind_corner_attr(f, c, a, read_one(m.attrs.binding_reg_cornerlist(r, a), a, m.conn.org(f, c), true));
	}
	prog.end();

	b200::FlatMesh flat;
	b200::flatten(m, order, nullptr, flat);
	std::vector<const uint8_t *> tptr(nlists);
	std::vector<uint32_t> tcnt(nlists);
	for (size_t l = 0; l < nlists; ++l) { tptr[l] = types[l].empty() ? nullptr : types[l].data(); tcnt[l] = (uint32_t)types[l].size(); }
	flat.desc.emit_type = tptr.data();
	flat.desc.emit_count = tcnt.data();
	if (hb_attr_decode(b200::context(), &flat.desc) != 0) b200::fail("attr_decode");
}

}
}

#include "formats/hry/reader.cc"
