/*
 * ref_harness.cc -- thin C interface around the UNMODIFIED reference (compiled from
 * /root/reference by oracle/Makefile into oracle/_ref/libharry_ref.so).  TEST INFRASTRUCTURE:
 * it is the source of golden vectors, the pin for oracle/harry_oracle.c, and the timed CPU
 * baseline ("kind": "reference") of bench.py.  No reference source is copied here; the
 * reference's translation units are #included / linked where they lie.
 *
 * The harness plugs its own writer / reader types into the reference's template seams
 * attrcode::AttrCoder<WR> / AttrDecoder<RD> (formats/hry/attrcode.h:291-300, 420-435) to capture
 * or replay the symbol streams, exactly as SURVEY.md section 8c describes.
 */
#include <cstdint>
#include <chrono>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "formats/unified_reader.h"
#include "formats/unified_writer.h"
#include "structs/mesh.h"
#include "structs/quant.h"
/* pulls in hry::writer::{MeshHandle,HeaderWriter,compress,write} and hry::reader::{...} */
#include "formats/hry/writer.cc"
#include "formats/hry/reader.cc"

#include "../include/harry_b200.h"

namespace {

std::string g_err;

struct NullBuf : std::streambuf {
	int overflow(int c) override { return c; }
	std::streamsize xsputn(const char *, std::streamsize n) override { return n; }
};

/* captured symbol streams in the layout of hb_streams */
struct ListCapture {
	std::vector<uint8_t> type;
	std::vector<uint32_t> aux;
	std::vector<uint8_t> sym;
	uint32_t ndata = 0;
};
struct Capture {
	std::vector<uint16_t> reg_vtx, reg_face;
	std::vector<ListCapture> lists;
};

/* WR concept of AttrCoder (formats/hry/io.h:90-116) */
struct CaptureWriter {
	Capture &cap;
	explicit CaptureWriter(Capture &c) : cap(c) {}
	void attr_data(mixing::View e, mesh::listidx_t l)
	{
		ListCapture &lc = cap.lists[l];
		lc.type.push_back(hry::DATA);
		lc.aux.push_back(0);
		for (int i = 0; i < e.fmt.size(); ++i) lc.sym.insert(lc.sym.end(), e.data(i), e.data(i) + e.bytes(i));
		++lc.ndata;
	}
	void attr_ghist(uint32_t idx, mesh::listidx_t l) { cap.lists[l].type.push_back(hry::HIST); cap.lists[l].aux.push_back(idx); }
	void attr_lhist(uint16_t idx, mesh::listidx_t l) { cap.lists[l].type.push_back(hry::LHIST); cap.lists[l].aux.push_back(idx); }
	void reg_face(mesh::regidx_t r) { cap.reg_face.push_back(r); }
	void reg_vtx(mesh::regidx_t r) { cap.reg_vtx.push_back(r); }
};

/* WR that drops everything: times prediction + residual without the arithmetic coder */
struct NullWriter {
	uint64_t sink = 0;
	void attr_data(mixing::View e, mesh::listidx_t) { if (e.fmt.size()) sink += e.data(0)[0]; }
	void attr_ghist(uint32_t idx, mesh::listidx_t) { sink += idx; }
	void attr_lhist(uint16_t idx, mesh::listidx_t) { sink += idx; }
	void reg_face(mesh::regidx_t r) { sink += r; }
	void reg_vtx(mesh::regidx_t r) { sink += r; }
};

/* RD concept of AttrDecoder (formats/hry/io.h:207-230) that forwards to the real io::reader and
 * records what it returned */
struct LoggingReader {
	hry::io::reader &rd;
	Capture &cap;
	LoggingReader(hry::io::reader &r, Capture &c) : rd(r), cap(c) {}
	void attr_data(mixing::View e, mesh::listidx_t l)
	{
		rd.attr_data(e, l);
		ListCapture &lc = cap.lists[l];
		lc.aux.push_back(0);
		for (int i = 0; i < e.fmt.size(); ++i) lc.sym.insert(lc.sym.end(), e.data(i), e.data(i) + e.bytes(i));
		++lc.ndata;
	}
	hry::AttrType attr_type(mesh::listidx_t l) { hry::AttrType t = rd.attr_type(l); cap.lists[l].type.push_back(t); return t; }
	uint32_t attr_ghist(mesh::listidx_t l) { uint32_t v = rd.attr_ghist(l); cap.lists[l].aux.push_back(v); return v; }
	uint16_t attr_lhist(mesh::listidx_t l) { uint16_t v = rd.attr_lhist(l); cap.lists[l].aux.push_back(v); return v; }
	mesh::regidx_t reg_face() { mesh::regidx_t r = rd.reg_face(); cap.reg_face.push_back(r); return r; }
	mesh::regidx_t reg_vtx() { mesh::regidx_t r = rd.reg_vtx(); cap.reg_vtx.push_back(r); return r; }
};

/* RD that replays captured streams (no arithmetic decoder): times reconstruction only */
struct ReplayReader {
	const Capture &cap;
	std::vector<size_t> pe, ps; /* per list: emission cursor, symbol byte cursor */
	size_t pv = 0, pf = 0;
	explicit ReplayReader(const Capture &c) : cap(c), pe(c.lists.size(), 0), ps(c.lists.size(), 0) {}
	void attr_data(mixing::View e, mesh::listidx_t l)
	{
		const ListCapture &lc = cap.lists[l];
		for (int i = 0; i < e.fmt.size(); ++i) {
			std::memcpy(e.data(i), lc.sym.data() + ps[l], e.bytes(i));
			ps[l] += e.bytes(i);
		}
		++pe[l];
	}
	hry::AttrType attr_type(mesh::listidx_t l) { return (hry::AttrType)cap.lists[l].type[pe[l]]; }
	uint32_t attr_ghist(mesh::listidx_t l) { return cap.lists[l].aux[pe[l]++]; }
	uint16_t attr_lhist(mesh::listidx_t l) { return (uint16_t)cap.lists[l].aux[pe[l]++]; }
	mesh::regidx_t reg_face() { return cap.reg_face[pf++]; }
	mesh::regidx_t reg_vtx() { return cap.reg_vtx[pv++]; }
};

hb_streams *to_streams(const Capture &cap, const mesh::Mesh &mesh, const std::vector<std::vector<uint64_t>> *hists)
{
	hb_streams *s = (hb_streams *)std::calloc(1, sizeof(hb_streams));
	s->n_vtx = (uint32_t)cap.reg_vtx.size();
	s->n_face = (uint32_t)cap.reg_face.size();
	s->reg_vtx = (uint16_t *)std::malloc(sizeof(uint16_t) * (cap.reg_vtx.size() + 1));
	s->reg_face = (uint16_t *)std::malloc(sizeof(uint16_t) * (cap.reg_face.size() + 1));
	std::memcpy(s->reg_vtx, cap.reg_vtx.data(), sizeof(uint16_t) * cap.reg_vtx.size());
	std::memcpy(s->reg_face, cap.reg_face.data(), sizeof(uint16_t) * cap.reg_face.size());
	s->nlists = (uint16_t)cap.lists.size();
	s->lists = (hb_list_streams *)std::calloc(cap.lists.size() + 1, sizeof(hb_list_streams));
	for (size_t l = 0; l < cap.lists.size(); ++l) {
		const ListCapture &lc = cap.lists[l];
		hb_list_streams &ls = s->lists[l];
		const mixing::Fmt &fmt = const_cast<mesh::Mesh &>(mesh).attrs[l].fmt();
		uint32_t stride = 0;
		for (int j = 0; j < fmt.size(); ++j) stride += mixing::SIZES[fmt.stype(j)];
		ls.n_emit = (uint32_t)lc.type.size();
		ls.n_data = lc.ndata;
		ls.sym_stride = stride;
		ls.type = (uint8_t *)std::malloc(lc.type.size() + 1);
		ls.aux = (uint32_t *)std::malloc(sizeof(uint32_t) * (lc.aux.size() + 1));
		ls.symbols = (uint8_t *)std::malloc(lc.sym.size() + 1);
		ls.hist = (uint64_t *)std::calloc((size_t)stride * 256 + 1, sizeof(uint64_t));
		std::memcpy(ls.type, lc.type.data(), lc.type.size());
		std::memcpy(ls.aux, lc.aux.data(), sizeof(uint32_t) * lc.aux.size());
		std::memcpy(ls.symbols, lc.sym.data(), lc.sym.size());
		if (hists && l < hists->size() && (*hists)[l].size() == (size_t)stride * 256)
			std::memcpy(ls.hist, (*hists)[l].data(), sizeof(uint64_t) * stride * 256);
		for (size_t k = 0; k < lc.type.size(); ++k) ls.type_hist[lc.type[k]]++;
	}
	return s;
}

/* final adaptive frequency tables of the real models: C[s] - 1 per (component, byte) context
 * (arith/model.h:36-47, arith/stat_adaptive.h:36; ModelVector type map formats/hry/models.h:129-145) */
template <typename T>
void grab_hist(arith::Model<uint64_t> *mdl, std::vector<uint64_t> &out)
{
	auto *mm = static_cast<arith::ModelMult<T, arith::AdaptiveStatisticsModule<>> *>(mdl);
	for (size_t b = 0; b < sizeof(T); ++b)
		for (int s = 0; s < 256; ++s) out.push_back(mm->stats[b].C[s] - 1);
}
void model_hists(hry::HryModels &models, mesh::Mesh &mesh, std::vector<std::vector<uint64_t>> &out)
{
	out.assign(mesh.attrs.size(), std::vector<uint64_t>());
	for (size_t l = 0; l < mesh.attrs.size(); ++l) {
		const mixing::Fmt &fmt = mesh.attrs[l].fmt();
		for (int j = 0; j < fmt.size(); ++j) {
			arith::Model<uint64_t> *mdl = (*models.attr_data[l])[j];
			switch (fmt.stype(j)) {
			case mixing::FLOAT: grab_hist<uint32_t>(mdl, out[l]); break;
			case mixing::DOUBLE: grab_hist<uint64_t>(mdl, out[l]); break;
			case mixing::ULONG: grab_hist<uint64_t>(mdl, out[l]); break;
			case mixing::LONG: grab_hist<int64_t>(mdl, out[l]); break;
			case mixing::UINT: grab_hist<uint32_t>(mdl, out[l]); break;
			case mixing::INT: grab_hist<int32_t>(mdl, out[l]); break;
			case mixing::USHORT: grab_hist<int16_t>(mdl, out[l]); break;
			case mixing::SHORT: grab_hist<uint16_t>(mdl, out[l]); break;
			case mixing::UCHAR: grab_hist<int8_t>(mdl, out[l]); break;
			case mixing::CHAR: grab_hist<uint8_t>(mdl, out[l]); break;
			default: break;
			}
		}
	}
}

double now_s()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

struct ref_mesh {
	mesh::Mesh mesh;
	std::vector<mesh::conn::fepair> order, order_f;
	bool traversed = false;
	Capture logged; /* streams logged while reading a .hry (decode side) */
	bool have_logged = false;
	std::vector<hb_list_desc> lists;
	std::vector<int32_t> off_face, off_corner, off_vtx;
	std::vector<std::vector<uint8_t>> groups;
	/* ref_snapshot / ref_restore */
	std::vector<std::vector<unsigned char>> snap_rows, snap_bounds;
	std::vector<mixing::Fmt> snap_fmt;
};

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

void ref_free(ref_mesh *m) { delete m; }

/* unified::reader::read (formats/unified_reader.h:78); for .hry inputs the attribute decoder runs
 * through a LoggingReader so that the residual rows / type / history symbols it consumed are kept. */
ref_mesh *ref_read(const char *path)
{
	ref_mesh *rm = new ref_mesh();
	try {
		std::string fn(path);
		std::ifstream is(fn, std::ifstream::binary);
		if (!is) throw std::runtime_error("cannot open " + fn);
		if (unified::reader::get_mesh_type(is, fn) == unified::reader::HRY) {
			/* same steps as hry::reader::read (formats/hry/reader.cc:179-193) */
			mesh::Builder builder(rm->mesh);
			hry::reader::HeaderReader hr(is);
			hr.read_syntax(builder);
			arith::Decoder<> coder(is);
			hry::HryModels models(builder.mesh);
			hry::io::reader rd(models, coder);
			rm->logged.lists.resize(rm->mesh.attrs.size());
			LoggingReader lrd(rd, rm->logged);
			hry::attrcode::AttrDecoder<LoggingReader> ac(builder, lrd);
			/* connectivity symbols are read through the plain reader */
			hry::reader::MeshHandle mh(rm->mesh);
			cbm::decode<hry::reader::MeshHandle, hry::io::reader, hry::attrcode::AttrDecoder<LoggingReader>, mesh::vtxidx_t, mesh::faceidx_t>(mh, rd, ac);
			progress::voidhandle prog;
			ac.decode(prog);
			rm->order = ac.order;
			rm->order_f.clear();
			for (mesh::faceidx_t f = 0; f < rm->mesh.num_face(); ++f) rm->order_f.push_back(mesh::conn::fepair(f, 0));
			rm->traversed = true;
			rm->have_logged = true;
		} else {
			unified::reader::read(is, fn, rm->mesh);
		}
	} catch (const std::exception &e) {
		g_err = e.what();
		delete rm;
		return nullptr;
	}
	return rm;
}

int ref_write(ref_mesh *rm, const char *path)
{
	try {
		unified::writer::write(std::string(path), rm->mesh);
	} catch (const std::exception &e) {
		g_err = e.what();
		return -1;
	}
	return 0;
}

int ref_set_bounds(ref_mesh *rm)
{
	quant::set_bounds(rm->mesh.attrs);
	return 0;
}

/* quant::requant(Attrs&, vector<Quant>, clear) (structs/quant.h:222-242); loq = n triples (l, o, q),
 * o == -1 selects every component like main.cc:80-84 */
int ref_requant(ref_mesh *rm, int n, const int *loq, int clear)
{
	try {
		std::vector<quant::Quant> q;
		for (int i = 0; i < n; ++i) {
			int l = loq[3 * i], o = loq[3 * i + 1], b = loq[3 * i + 2];
			if (l < 0 || l >= (int)rm->mesh.attrs.size()) throw std::runtime_error("Invalid list index");
			if (o == -1) {
				for (int k = 0; k < rm->mesh.attrs[l].fmt().size(); ++k) q.push_back(quant::Quant(l, k, b));
			} else {
				q.push_back(quant::Quant(l, o, b));
			}
		}
		quant::requant(rm->mesh.attrs, q, clear != 0);
	} catch (const std::exception &e) {
		g_err = e.what();
		return -1;
	}
	return 0;
}

/* cbm::encode with the real io::writer into a null sink (formats/hry/writer.cc:200-211): records
 * the traversal order and leaves the twin table as the attribute coder will see it. */
int ref_traverse(ref_mesh *rm)
{
	try {
		NullBuf nb;
		std::ostream os(&nb);
		arith::Encoder<> coder(os);
		hry::HryModels models(rm->mesh);
		hry::io::writer wr(models, coder);
		hry::attrcode::AttrCoder<hry::io::writer> ac(rm->mesh, wr);
		hry::writer::MeshHandle mh(rm->mesh);
		cbm::encode<hry::writer::MeshHandle, hry::io::writer, hry::attrcode::AttrCoder<hry::io::writer>, mesh::vtxidx_t, mesh::faceidx_t>(mh, wr, ac);
		rm->order = ac.order;
		rm->order_f = ac.order_f;
		rm->traversed = true;
	} catch (const std::exception &e) {
		g_err = e.what();
		return -1;
	}
	return 0;
}

/* Flatten the mesh into an hb_mesh_desc whose pointers alias the mesh's own vectors. */
int ref_desc(ref_mesh *rm, hb_mesh_desc *d)
{
	mesh::Mesh &m = rm->mesh;
	std::memset(d, 0, sizeof *d);
	d->nv = m.num_vtx();
	d->nf = m.num_face();
	d->ne = m.num_edge();
	d->edges = m.conn.edges.data();
	d->face_off = m.faces.offsets.data();
	d->order = rm->order.data();
	d->norder = (uint32_t)rm->order.size();
	d->order_f = rm->order_f.data();
	d->norder_f = (uint32_t)rm->order_f.size();
	d->vtx_regs = m.attrs.vtx_regs.data();
	d->face_regs = m.attrs.face_regs.data();
	d->nb_face = m.attrs.num_bindings_face;
	d->nb_vtx = m.attrs.num_bindings_vtx;
	d->nb_corner = m.attrs.num_bindings_corner;
	d->nregs_face = m.attrs.num_regs_face();
	d->nregs_vtx = m.attrs.num_regs_vtx();
	d->nlists = (uint16_t)m.attrs.size();
	d->bind_face_attr = m.attrs.bindings_face_attr.data();
	d->bind_vtx_attr = m.attrs.bindings_vtx_attr.data();
	d->bind_corner_attr = m.attrs.bindings_corner_attr.data();
	rm->off_face.assign(m.attrs.off_reg_facelist.begin(), m.attrs.off_reg_facelist.end());
	rm->off_corner.assign(m.attrs.off_reg_cornerlist.begin(), m.attrs.off_reg_cornerlist.end());
	rm->off_vtx.assign(m.attrs.off_reg_vtxlist.begin(), m.attrs.off_reg_vtxlist.end());
	d->off_reg_face = rm->off_face.data();
	d->off_reg_corner = rm->off_corner.data();
	d->off_reg_vtx = rm->off_vtx.data();
	d->reg_facelist = m.attrs.bindings_reg_facelist.data();
	d->reg_cornerlist = m.attrs.bindings_reg_cornerlist.data();
	d->reg_vtxlist = m.attrs.bindings_reg_vtxlist.data();
	rm->lists.assign(m.attrs.size(), hb_list_desc());
	rm->groups.assign(m.attrs.size(), std::vector<uint8_t>());
	for (size_t l = 0; l < m.attrs.size(); ++l) {
		hb_list_desc &L = rm->lists[l];
		std::memset(&L, 0, sizeof L);
		const mixing::Fmt &fmt = m.attrs[l].fmt();
		if (fmt.size() > HB_MAX_COMP) { g_err = "too many components"; return -2; }
		L.rows = m.attrs[l].data();
		L.nrows = (uint32_t)m.attrs[l].size();
		L.stride = fmt.bytes();
		L.ncomp = (uint16_t)fmt.size();
		L.target = (uint8_t)m.attrs[l].target;
		for (int j = 0; j < fmt.size(); ++j) {
			L.type[j] = (uint8_t)fmt.type(j);
			L.quant[j] = (uint8_t)fmt.quant(j);
			L.offset[j] = (uint16_t)fmt.offset(j);
		}
		/* interpretation-group leader of every component (structs/quant.h:54-59) */
		rm->groups[l].assign(fmt.size(), 0);
		for (int j = 0; j < fmt.size(); ++j) rm->groups[l][j] = (uint8_t)j;
		mixing::Interps &ip = m.attrs[l].interps();
		for (int i = 0; i < ip.size(); ++i)
			for (int j = 0; j < ip.len(i); ++j)
				if (ip.off(i) >= 0 && ip.off(i) + j < fmt.size()) rm->groups[l][ip.off(i) + j] = (uint8_t)ip.off(i);
	}
	d->lists = rm->lists.data();
	return 0;
}

const uint8_t *ref_groups(ref_mesh *rm, int l) { return rm->groups[l].data(); }
/* bounds rows of list l: which = 0 min, 1 max, 2 scale (structs/attr.h:78-93) */
const void *ref_bounds_row(ref_mesh *rm, int l, int which) { return rm->mesh.attrs[l].bounds()[which].data(); }
int ref_set_scale(ref_mesh *rm, int l) { quant::set_scale(rm->mesh.attrs[l]); return 0; }
int ref_has_logged(ref_mesh *rm) { return rm->have_logged ? 1 : 0; }

/* The real attribute coder with a capturing writer + the real models' final histograms. */
int ref_attr_encode(ref_mesh *rm, hb_streams **out)
{
	try {
		if (!rm->traversed) throw std::runtime_error("ref_attr_encode: call ref_traverse first");
		Capture cap;
		cap.lists.resize(rm->mesh.attrs.size());
		CaptureWriter cw(cap);
		hry::attrcode::AttrCoder<CaptureWriter> ac(rm->mesh, cw);
		ac.order = rm->order;
		ac.order_f = rm->order_f;
		progress::voidhandle prog;
		ac.encode(prog);
		/* histograms from the real io::writer + arithmetic coder run */
		std::vector<std::vector<uint64_t>> hists;
		{
			NullBuf nb;
			std::ostream os(&nb);
			arith::Encoder<> coder(os);
			hry::HryModels models(rm->mesh);
			hry::io::writer wr(models, coder);
			hry::attrcode::AttrCoder<hry::io::writer> ac2(rm->mesh, wr);
			ac2.order = rm->order;
			ac2.order_f = rm->order_f;
			ac2.encode(prog);
			coder.flush();
			model_hists(models, rm->mesh, hists);
		}
		*out = to_streams(cap, rm->mesh, &hists);
	} catch (const std::exception &e) {
		g_err = e.what();
		return -1;
	}
	return 0;
}

/* streams logged by the real decoder while reading the .hry */
int ref_logged_streams(ref_mesh *rm, hb_streams **out)
{
	if (!rm->have_logged) { g_err = "mesh was not read from .hry"; return -1; }
	*out = to_streams(rm->logged, rm->mesh, nullptr);
	return 0;
}

void ref_streams_free(hb_streams *s)
{
	if (!s) return;
	for (int l = 0; l < s->nlists; ++l) { std::free(s->lists[l].type); std::free(s->lists[l].aux); std::free(s->lists[l].symbols); std::free(s->lists[l].hist); }
	std::free(s->lists); std::free(s->reg_vtx); std::free(s->reg_face); std::free(s);
}

/* ---- timed CPU baseline: the reference functions the CUDA kernels replace, one core ------- */
/* times[0] set_bounds, [1] requant(quantize), [2] AttrCoder<NullWriter>::encode (vertices only),
 * [3] AttrCoder<NullWriter>::encode (full: faces + corner fan walks as the reference does),
 * [4] AttrDecoder<ReplayReader>::decode, [5] requant(clear) */
int ref_time_path(ref_mesh *rm, int n, const int *loq, double *times)
{
	try {
		double t0 = now_s();
		quant::set_bounds(rm->mesh.attrs);
		double t1 = now_s();
		times[0] = t1 - t0;
		if (ref_requant(rm, n, loq, 0)) return -1;
		double t2 = now_s();
		times[1] = t2 - t1;
		if (!rm->traversed && ref_traverse(rm)) return -1;
		progress::voidhandle prog;
		{
			NullWriter nw;
			hry::attrcode::AttrCoder<NullWriter> ac(rm->mesh, nw);
			ac.order = rm->order;
			double a = now_s();
			ac.encode(prog);
			times[2] = now_s() - a;
		}
		{
			NullWriter nw;
			hry::attrcode::AttrCoder<NullWriter> ac(rm->mesh, nw);
			ac.order = rm->order;
			ac.order_f = rm->order_f;
			double a = now_s();
			ac.encode(prog);
			times[3] = now_s() - a;
		}
		times[4] = 0;
		times[5] = 0;
	} catch (const std::exception &e) {
		g_err = e.what();
		return -1;
	}
	return 0;
}

/* ---- several timed steps on ONE prepared mesh (bench.py --impl reference: prepare once, time K steps) ---- */
/* The timed functions work in place: keep / restore a copy of all attribute rows, bounds rows and formats. */
int ref_snapshot(ref_mesh *rm)
{
	rm->snap_rows.clear(); rm->snap_bounds.clear(); rm->snap_fmt.clear();
	for (size_t l = 0; l < rm->mesh.attrs.size(); ++l) {
		mesh::attr::Attr &a = rm->mesh.attrs[l];
		rm->snap_rows.emplace_back(a.data(), a.data() + a.bytes());
		rm->snap_bounds.emplace_back(a.bounds().data(), a.bounds().data() + a.bounds().bytes());
		rm->snap_fmt.push_back(a.fmt());
	}
	return 0;
}
int ref_restore(ref_mesh *rm)
{
	if (rm->snap_rows.size() != rm->mesh.attrs.size()) { g_err = "restore without snapshot"; return -1; }
	for (size_t l = 0; l < rm->mesh.attrs.size(); ++l) {
		mesh::attr::Attr &a = rm->mesh.attrs[l];
		a.tmp() = rm->snap_fmt[l];
		a.restore_fmt();
		if (a.bytes() != rm->snap_rows[l].size()) { g_err = "restore: list size changed"; return -1; }
		if (a.bytes()) std::memcpy(a.data(), rm->snap_rows[l].data(), a.bytes());
		if (a.bounds().bytes()) std::memcpy(a.bounds().data(), rm->snap_bounds[l].data(), a.bounds().bytes());
	}
	return 0;
}
/* encode side of one step: times[0] set_bounds, [1] requant(quantize), [3] AttrCoder<NullWriter>::encode (full);
 * the Cut-Border-Machine traversal (outside the path) runs once, untimed, the first time */
int ref_time_encode_step(ref_mesh *rm, int n, const int *loq, double *times)
{
	try {
		double t0 = now_s();
		quant::set_bounds(rm->mesh.attrs);
		double t1 = now_s();
		times[0] = t1 - t0;
		if (ref_requant(rm, n, loq, 0)) return -1;
		times[1] = now_s() - t1;
		if (!rm->traversed && ref_traverse(rm)) return -1;
		progress::voidhandle prog;
		NullWriter nw;
		hry::attrcode::AttrCoder<NullWriter> ac(rm->mesh, nw);
		ac.order = rm->order;
		ac.order_f = rm->order_f;
		double a = now_s();
		ac.encode(prog);
		times[3] = now_s() - a;
		times[2] = 0;
	} catch (const std::exception &e) {
		g_err = e.what();
		return -1;
	}
	return 0;
}

/* decode side: `hry_path` is read with the real decoder once (logging); every step then times (a) the real
 * AttrDecoder fed from the logged streams on a fresh mesh with the same connectivity and (b)
 * quant::requant(clear).  times[4], times[5] as above. */
struct ref_decode_ctx {
	ref_mesh *src;
	std::string path;
};
ref_decode_ctx *ref_decode_open(const char *hry_path)
{
	ref_mesh *src = ref_read(hry_path);
	if (!src) return nullptr;
	ref_decode_ctx *c = new ref_decode_ctx();
	c->src = src;
	c->path = hry_path;
	return c;
}
void ref_decode_close(ref_decode_ctx *c)
{
	if (!c) return;
	ref_free(c->src);
	delete c;
}
int ref_decode_step(ref_decode_ctx *c, double *times)
{
	ref_mesh *src = c->src;
	int rc = 0;
	try {
		/* fresh mesh: header again, connectivity copied, attribute decode replayed */
		ref_mesh dst;
		std::ifstream is(c->path, std::ifstream::binary);
		mesh::Builder builder(dst.mesh);
		hry::reader::HeaderReader hr(is);
		hr.read_syntax(builder);
		dst.mesh.conn.edges = src->mesh.conn.edges;
		dst.mesh.conn.mnum_vtx = src->mesh.conn.mnum_vtx;
		dst.mesh.conn.mnum_tri = src->mesh.conn.mnum_tri;
		dst.mesh.faces.offsets = src->mesh.faces.offsets;
		ReplayReader rr(src->logged);
		hry::attrcode::AttrDecoder<ReplayReader> ac(builder, rr);
		ac.order = src->order;
		progress::voidhandle prog;
		double a = now_s();
		ac.decode(prog);
		times[4] = now_s() - a;
		/* sanity: identical to what the real decoder produced */
		for (size_t l = 0; l < dst.mesh.attrs.size(); ++l) {
			if (dst.mesh.attrs[l].bytes() != src->mesh.attrs[l].bytes() ||
			    std::memcmp(dst.mesh.attrs[l].data(), src->mesh.attrs[l].data(), dst.mesh.attrs[l].bytes()) != 0)
				throw std::runtime_error("replayed decode differs from the real decode");
		}
		a = now_s();
		quant::requant(dst.mesh.attrs, std::vector<quant::Quant>(), true);
		times[5] = now_s() - a;
	} catch (const std::exception &e) {
		g_err = e.what();
		rc = -1;
	}
	return rc;
}
int ref_time_decode(const char *hry_path, double *times)
{
	ref_decode_ctx *c = ref_decode_open(hry_path);
	if (!c) return -1;
	const int rc = ref_decode_step(c, times);
	ref_decode_close(c);
	return rc;
}

/* The reference's own twin matching on a face list: conn::Builder (structs/conn.h:172-233) driven the way the
 * readers drive it -- face_begin, set_org per corner, face_end (formats/ply/reader.cc:349-353).  edges_out receives
 * Conn::edges (12-byte edgeorg records), *seconds the time of the loop alone. */
int ref_twin_match(uint32_t nf, const uint32_t *face_off, const uint32_t *org, void *edges_out, double *seconds)
{
	try {
		static_assert(sizeof(mesh::conn::Conn::edgeorg) == 12, "Conn::edgeorg layout");
		mesh::Faces faces;
		mesh::conn::Conn conn(faces);
		mesh::conn::Builder b(conn);
		b.reserve(nf);
		double t0 = now_s();
		for (uint32_t f = 0; f < nf; ++f) {
			const uint32_t n = face_off[f + 1] - face_off[f];
			b.face_begin((mesh::ledgeidx_t)n);
			for (uint32_t e = 0; e < n; ++e) b.set_org(org[face_off[f] + e]);
			b.face_end();
		}
		double t1 = now_s();
		if (seconds) *seconds = t1 - t0;
		if (edges_out && !conn.edges.empty()) std::memcpy(edges_out, conn.edges.data(), conn.edges.size() * 12);
		return 0;
	} catch (std::exception &e) {
		g_err = e.what();
		return -1;
	}
}

} // extern "C"
