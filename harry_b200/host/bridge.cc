// bridge.cc -- see bridge.h
#include "bridge.h"

#include <cstdlib>
#include <cstring>
#include <future>

namespace b200 {

static hb_ctx *g_ctx = nullptr;

// Creating the CUDA context takes 1-2 s in a fresh process -- as long as the whole attribute path of a 10M-vertex
// mesh.  It starts on a helper thread when the program starts and runs under the file parsing (main.cc reads the
// input first); the first GPU call joins it.  A failure (no device) surfaces at that first call, as before.
struct Warmup {
	std::future<int> done;
	std::string error;
	Warmup()
	{
		done = std::async(std::launch::async, [this]() -> int {
			const char *dev = std::getenv("HARRY_B200_DEVICE");
			const int rc = hb_ctx_create(dev ? std::atoi(dev) : 0, &g_ctx);
			if (rc != 0) error = hb_last_error(nullptr);
			// main.cc never touches the rows between set_bounds -> requant -> write / read -> requant(clear): their device
			// copy stays where it is between the calls (one upload per pipeline instead of one per call)
			else hb_ctx_set_row_cache(g_ctx, 1);
			return rc;
		});
	}
};
static Warmup g_warmup;

hb_ctx *context()
{
	if (g_warmup.done.valid() && g_warmup.done.get() != 0) {
		g_ctx = nullptr;
		throw std::runtime_error(std::string("harry_b200: ") + g_warmup.error);
	}
	if (!g_ctx) throw std::runtime_error(std::string("harry_b200: ") + (g_warmup.error.empty() ? "no context" : g_warmup.error));
	return g_ctx;
}

void fail(const char *what)
{
	throw std::runtime_error(std::string("harry_b200 ") + what + ": " + hb_last_error(g_ctx));
}

hb_list_desc describe_list(mesh::attr::Attr &attr)
{
	hb_list_desc L;
	std::memset(&L, 0, sizeof L);
	const mixing::Fmt &fmt = attr.fmt();
	if (fmt.size() > HB_MAX_COMP) throw std::runtime_error("harry_b200: attribute list with more than 32 components");
	L.rows = attr.data();
	L.nrows = (uint32_t)attr.size();
	L.stride = (uint32_t)fmt.bytes();
	L.ncomp = (uint16_t)fmt.size();
	L.target = (uint8_t)attr.target;
	for (int j = 0; j < fmt.size(); ++j) {
		L.type[j] = (uint8_t)fmt.type(j);
		L.quant[j] = (uint8_t)fmt.quant(j);
		L.offset[j] = (uint16_t)fmt.offset(j);
	}
	return L;
}

void flatten(mesh::Mesh &m, const std::vector<mesh::conn::fepair> &order, const std::vector<mesh::conn::fepair> *order_f, FlatMesh &out)
{
	static_assert(sizeof(mesh::conn::Conn::edgeorg) == 12, "Conn::edgeorg layout (structs/conn.h:73-76)");
	static_assert(sizeof(mesh::conn::fepair) == 8, "fepair layout (structs/conn.h:22-31)");
	hb_mesh_desc &d = out.desc;
	std::memset(&d, 0, sizeof d);
	d.nv = m.num_vtx();
	d.nf = m.num_face();
	d.ne = m.num_edge();
	d.edges = m.conn.edges.data();
	d.face_off = m.faces.offsets.data();
	d.order = order.data();
	d.norder = (uint32_t)order.size();
	d.order_f = order_f ? order_f->data() : nullptr;
	d.norder_f = order_f ? (uint32_t)order_f->size() : 0;
	d.vtx_regs = m.attrs.vtx_regs.data();
	d.face_regs = m.attrs.face_regs.data();
	d.nb_face = m.attrs.num_bindings_face;
	d.nb_vtx = m.attrs.num_bindings_vtx;
	d.nb_corner = m.attrs.num_bindings_corner;
	d.nregs_face = m.attrs.num_regs_face();
	d.nregs_vtx = m.attrs.num_regs_vtx();
	d.nlists = (uint16_t)m.attrs.size();
	d.bind_face_attr = m.attrs.bindings_face_attr.data();
	d.bind_vtx_attr = m.attrs.bindings_vtx_attr.data();
	d.bind_corner_attr = m.attrs.bindings_corner_attr.data();
	out.off_face.assign(m.attrs.off_reg_facelist.begin(), m.attrs.off_reg_facelist.end());
	out.off_corner.assign(m.attrs.off_reg_cornerlist.begin(), m.attrs.off_reg_cornerlist.end());
	out.off_vtx.assign(m.attrs.off_reg_vtxlist.begin(), m.attrs.off_reg_vtxlist.end());
	d.off_reg_face = out.off_face.data();
	d.off_reg_corner = out.off_corner.data();
	d.off_reg_vtx = out.off_vtx.data();
	d.reg_facelist = m.attrs.bindings_reg_facelist.data();
	d.reg_cornerlist = m.attrs.bindings_reg_cornerlist.data();
	d.reg_vtxlist = m.attrs.bindings_reg_vtxlist.data();
	out.lists.clear();
	for (size_t l = 0; l < m.attrs.size(); ++l) out.lists.push_back(describe_list(m.attrs[l]));
	d.lists = out.lists.data();
}

}

namespace quant {

// quant::set_bounds (structs/quant.h:39-44): min / max rows of every list, on the GPU
void set_bounds_b200(mesh::attr::Attrs &attrs)
{
	hb_ctx *ctx = b200::context();
	for (mesh::listidx_t i = 0; i < attrs.size(); ++i) {
		mesh::attr::Attr &attr = attrs[i];
		if (attr.fmt().size() == 0) continue;
		hb_list_desc L = b200::describe_list(attr);
		if (hb_bounds(ctx, &L, attr.min().data(), attr.max().data()) != 0) b200::fail("set_bounds");
	}
}

void finish_read_b200(mesh::Mesh &mesh)
{
	static_assert(sizeof(mesh::conn::Conn::edgeorg) == 12, "hb_twin_match works in place on 12-byte Conn::edgeorg records");
	const uint32_t nf = (uint32_t)mesh.faces.size();
	if (nf && hb_twin_match(b200::context(), (uint32_t)mesh.conn.num_vtx(), nf, mesh.faces.offsets.data(), mesh.conn.edges.data(), 12,
	                        mesh.conn.edges.data()) != 0)
		b200::fail("twin_match");
	set_bounds_b200(mesh.attrs);
}

// quant::requant(Attrs&, vector<Quant>, clear) (structs/quant.h:222-242): the format bookkeeping
// (backup_fmt / tmp / restore_fmt) and set_scale stay the reference's; the row conversion
// requant(Attr&, Fmt) (:215-221) runs on the GPU
void requant_b200(mesh::attr::Attrs &attrs, const std::vector<Quant> &quant, bool clear)
{
	hb_ctx *ctx = b200::context();
	for (mesh::listidx_t l = 0; l < attrs.size(); ++l) attrs[l].backup_fmt();
	if (clear)
		for (mesh::listidx_t l = 0; l < attrs.size(); ++l)
			for (int i = 0; i < attrs[l].fmt().size(); ++i) attrs[l].tmp().setquant(i, 0);
	for (size_t i = 0; i < quant.size(); ++i) attrs[quant[i].l].tmp().setquant(quant[i].o, quant[i].q);
	for (mesh::listidx_t l = 0; l < attrs.size(); ++l) {
		mesh::attr::Attr &attr = attrs[l];
		set_scale(attr);
		if (attr.fmt().size()) {
			hb_list_desc L = b200::describe_list(attr);
			uint8_t nq[HB_MAX_COMP] = { 0 };
			for (int j = 0; j < attr.tmp().size(); ++j) nq[j] = (uint8_t)attr.tmp().quant(j);
			if (hb_requant(ctx, &L, nq, attr.min().data(), attr.scale().data()) != 0) b200::fail("requant");
		}
		attr.restore_fmt();
	}
}

}
