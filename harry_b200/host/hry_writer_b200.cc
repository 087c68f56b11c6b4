// hry_writer_b200.cc -- the reference .hry writer (formats/hry/writer.cc, compiled unchanged) with
// the attribute stage AttrCoder<io::writer>::encode (attrcode.h:396-416, call site writer.cc:212)
// replaced by the GPU path: an explicit specialization of that member template, declared before
// writer.cc is parsed, so compress() binds to it.  The symbol streams computed on the GPU are
// replayed into the unchanged io::writer in the reference's emission order, so the adaptive
// arithmetic coder produces byte-identical output.
#include <cstdint>
#include <cstddef>
#include <cstring>
#include "formats/hry/writer.h"
#include "formats/hry/common.h"
#include "formats/hry/attrcode.h"
#include "formats/hry/io.h"
#include "utils/progress.h"
#include "bridge.h"

namespace hry {
namespace attrcode {

template <>
template <>
void AttrCoder<io::writer>::encode<progress::handle>(progress::handle &prog)
{
	b200::FlatMesh flat;
	b200::flatten(mesh, order, &order_f, flat);
	hb_streams *s = nullptr;
	if (hb_attr_encode(b200::context(), &flat.desc, &s) != 0) b200::fail("attr_encode");

	std::vector<uint32_t> emit(mesh.attrs.size(), 0), data(mesh.attrs.size(), 0);
	std::vector<mixing::Fmt> fmts;
	std::vector<std::vector<unsigned char>> rowbuf(mesh.attrs.size());
	for (mesh::listidx_t l = 0; l < mesh.attrs.size(); ++l) rowbuf[l].resize(mesh.attrs[l].fmt().bytes() + 8);
	// one emission of list l: type symbol, then residual row or history offset (io.h:90-108)
	auto emit_one = [&](mesh::listidx_t l) {
		const hb_list_streams &ls = s->lists[l];
		const uint32_t k = emit[l]++;
		switch (ls.type ? ls.type[k] : (uint8_t)HB_DATA) { // NULL: every emission is a DATA row
		case HB_DATA: {
			const mixing::Fmt &fmt = mesh.attrs[l].fmt();
			const uint8_t *src = ls.symbols + (size_t)data[l]++ * ls.sym_stride;
			mixing::View row(rowbuf[l].data(), fmt);
			for (int j = 0; j < fmt.size(); ++j) {
				std::memcpy(row.data(j), src, row.bytes(j));
				src += row.bytes(j);
			}
			wr.attr_data(row, l);
			break;
		}
		case HB_HIST: wr.attr_ghist(ls.aux[k], l); break;
		default: wr.attr_lhist((uint16_t)ls.aux[k], l); break;
		}
	};

	prog.start(order.size());
	for (size_t i = 0; i < order.size(); ++i) { // attrcode.h:399-404, vtx_post :321-344
		const mesh::regidx_t r = s->reg_vtx ? s->reg_vtx[i] : 0; // NULL: single region
		wr.reg_vtx(r);
		for (mesh::listidx_t a = 0; a < mesh.attrs.num_bindings_vtx_reg(r); ++a) emit_one(mesh.attrs.binding_reg_vtxlist(r, a));
		prog(i);
	}
	for (size_t i = 0; i < order_f.size(); ++i) { // attrcode.h:405-414, face_post :345-365, corner_post :367-393
		const mesh::regidx_t r = s->reg_face ? s->reg_face[i] : 0;
		wr.reg_face(r);
		for (mesh::listidx_t a = 0; a < mesh.attrs.num_bindings_face_reg(r); ++a) emit_one(mesh.attrs.binding_reg_facelist(r, a));
		const int ne = mesh.conn.num_edges(order_f[i].f());
		const mesh::listidx_t nc = mesh.attrs.num_bindings_corner_reg(r);
		for (int c = 0; c < ne; ++c)
			for (mesh::listidx_t a = 0; a < nc; ++a) emit_one(mesh.attrs.binding_reg_cornerlist(r, a));
	}
	prog.end();
	hb_streams_free(s);
}

}
}

#include "formats/hry/writer.cc"
