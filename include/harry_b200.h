/*
 * harry_b200.h -- C ABI of the B200-native attribute path of the Harry mesh compressor.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ and no
 * torch types.  Each entry point replaces one call of the reference (paths relative to the
 * reference tree):
 *
 *   hb_bounds        <- quant::set_bounds(Attr&)                 structs/quant.h:30-38
 *                       (call sites formats/ply/reader.cc:428, formats/obj/reader.cc:1253)
 *   hb_requant       <- quant::requant(Attr&, const Fmt&)        structs/quant.h:114-221
 *                       (driver quant::requant(Attrs&,...) :222-242, call site main.cc:108)
 *   hb_attr_encode   <- attrcode::AttrCoder<WR>::encode(P&)      formats/hry/attrcode.h:396-416
 *                       (call site formats/hry/writer.cc:212)
 *   hb_attr_decode   <- attrcode::AttrDecoder<RD>::decode(P&)    formats/hry/attrcode.h:534-550
 *                       (call site formats/hry/reader.cc:192), after the host drained the
 *                       arithmetic-coded symbol stream into residual rows + binding tables
 *   hb_twin_match    <- mesh::Builder::add_edge / face_end / set_org  structs/conn.h:164-214
 *                       (the twin hash join the readers run per face corner, formats/ply/reader.cc:349-353,
 *                       formats/obj/reader.rl; SURVEY.md section 8f row f2 -- the step in front of the path)
 *
 * The same structs are consumed by the CPU oracle (oracle/harry_oracle.h: ho_* functions with
 * identical signatures), which is test infrastructure only.
 *
 * Conventions: every function returns 0 on success and a negative hb_status on failure;
 * hb_last_error() gives a message.  The library never frees or keeps caller memory beyond the
 * call unless stated.  An hb_ctx is single-owner (not thread safe); use one per host thread/GPU.
 * There is no CPU fallback: a missing GPU or a CUDA failure is an error.
 */
#ifndef HARRY_B200_H
#define HARRY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_MAX_COMP 32 /* components per attribute list */

typedef enum hb_status {
	HB_OK = 0,
	HB_ERR_CUDA = -1,        /* CUDA runtime failure (no device, OOM, launch error) */
	HB_ERR_INVALID = -2,     /* malformed descriptor */
	HB_ERR_UNSUPPORTED = -3, /* type/quant combination the reference itself leaves undefined */
	HB_ERR_NOMEM = -4
} hb_status;

/* mixing::Type, structs/mixing.h:19 (the enum VALUE is stored in .hry headers, writer.cc:163). */
typedef enum hb_type {
	HB_FLOAT = 0, HB_DOUBLE = 1, HB_ULONG = 2, HB_LONG = 3, HB_UINT = 4, HB_INT = 5,
	HB_USHORT = 6, HB_SHORT = 7, HB_UCHAR = 8, HB_CHAR = 9, HB_TYPE_NONE = 10
} hb_type;

/* mesh::attr::Target, structs/attr.h:22 */
typedef enum hb_target { HB_FACE = 0, HB_VTX = 1, HB_CORNER = 2, HB_TARGET_NONE = 3 } hb_target;

/* hry::AttrType, formats/hry/models.h:21 */
typedef enum hb_attr_type { HB_DATA = 0, HB_HIST = 1, HB_LHIST = 2 } hb_attr_type;

/*
 * One attribute list = mixing::Array + mixing::Fmt (structs/mixing.h:41-137, 378-441):
 * `nrows` AoS rows of `stride` bytes; component j of unquantized type `type[j]` lives at byte
 * `offset[j]`; when `quant[j] != 0` the slot holds the fixed-point value in the storage type
 * Fmt::quant_type(quant[j]) (u8 / u16 / u32 / u64) in its low bytes and the remaining bytes of
 * the slot are ignored on input and preserved on output.
 */
typedef struct hb_list_desc {
	void *rows;
	uint32_t nrows;
	uint32_t stride;
	uint16_t ncomp;
	uint8_t target; /* hb_target */
	uint8_t reserved;
	uint8_t type[HB_MAX_COMP];
	uint8_t quant[HB_MAX_COMP];
	uint16_t offset[HB_MAX_COMP];
} hb_list_desc;

/*
 * Flattened mesh::Mesh as the attribute coder sees it after the Cut-Border-Machine traversal.
 *
 * edges     ne records of 12 bytes, the raw mesh::conn::Conn::edgeorg layout (structs/conn.h:73-76):
 *           { u32 org; u32 twin_face; u16 twin_edge; u16 pad }.  Half-edge (f, e) has the global
 *           index face_off[f] + e (conn.h:123-126).  The twin table must be the one AFTER
 *           cbm::encode / cbm::decode mutated it (cbm/encoder.h:150,193-198).
 * face_off  nf + 1 CSR offsets (mesh::Faces::offsets, structs/faces.h:19).
 * order     traversal order of the vertices, norder records of 8 bytes = mesh::conn::fepair
 *           { u32 face; u16 edge; u16 pad } (attrcode.h:297,310-314); org(order[i]) is the vertex.
 * order_f   traversal order of the faces with their gate corner (attrcode.h:298,315-319).
 *           NULL means faces in index order with gate corner 0 (the decoder's order, :543-548).
 * *_regs, bind_*, off_reg_*, reg_*list: mesh::attr::Bindings (structs/attr.h:101-189).
 *
 * Decode-side value semantics (attrcode.h:443-531): attribute rows start zeroed and are filled in
 * DATA-emission order, so a prediction candidate whose row has not been emitted yet reads as 0.
 */
typedef struct hb_mesh_desc {
	uint32_t nv, nf, ne;
	const void *edges;
	const uint32_t *face_off;
	const void *order;
	uint32_t norder;
	const void *order_f;
	uint32_t norder_f;
	const uint16_t *vtx_regs;  /* nv */
	const uint16_t *face_regs; /* nf */
	uint16_t nb_face, nb_vtx, nb_corner; /* Bindings::num_bindings_* */
	uint16_t nregs_face, nregs_vtx;
	uint16_t nlists;
	const uint32_t *bind_face_attr;   /* nf * nb_face */
	const uint32_t *bind_vtx_attr;    /* nv * nb_vtx */
	const uint32_t *bind_corner_attr; /* ne * nb_corner */
	const int32_t *off_reg_face;   /* nregs_face + 1 */
	const int32_t *off_reg_corner; /* nregs_face + 1 */
	const int32_t *off_reg_vtx;    /* nregs_vtx + 1 */
	const uint16_t *reg_facelist, *reg_cornerlist, *reg_vtxlist;
	hb_list_desc *lists; /* nlists */
	/* Decode only, optional: emit_type[l] = the type symbols (hb_attr_type) of list l in emission
	 * order, as drained from the stream.  They identify which element carried each DATA row.
	 * NULL (or a NULL entry) = "the first element referencing a row carried it", which holds
	 * whenever no corner binding slot is shared by different lists across face regions (the
	 * reference's LocalHistory is keyed by slot, attrcode.h:296,377: there an LHIST hit can
	 * precede the DATA emission of the row it names). */
	const uint8_t *const *emit_type;
	const uint32_t *emit_count; /* nlists: length of emit_type[l] (required with emit_type) */
} hb_mesh_desc;

/*
 * Symbol streams of one attribute list, in the exact order the reference feeds its io::writer
 * (formats/hry/io.h:90-116): one `type` symbol per emission; for HIST / LHIST emissions the
 * history offset in `aux`; for DATA emissions one residual row in `symbols`.
 * A residual row is the byte-plane symbol sequence of arith::ModelMult (arith/model.h:57-66,
 * formats/hry/models.h:168-173): for component j = 0..ncomp-1, bytes 0..sizeof(stype_j)-1 of the
 * residual in little-endian order; byte position p of the row is coded with context (list, p).
 * hist[p * 256 + s] counts symbol s in context p  ( = AdaptiveStatisticsModule::C[s] - 1 after
 * coding, arith/stat_adaptive.h:36,77-82).
 */
typedef struct hb_list_streams {
	uint32_t n_emit;
	uint32_t n_data;
	uint32_t sym_stride; /* bytes per residual row == number of contexts */
	uint32_t reserved;
	uint8_t *type;     /* n_emit */
	uint32_t *aux;     /* n_emit */
	uint8_t *symbols;  /* n_data * sym_stride */
	uint64_t *hist;    /* sym_stride * 256 */
	uint64_t type_hist[4];
} hb_list_streams;

typedef struct hb_streams {
	uint32_t n_vtx;     /* == norder   */
	uint32_t n_face;    /* == norder_f */
	uint16_t *reg_vtx;  /* region symbol per traversed vertex (io.h:113) */
	uint16_t *reg_face; /* region symbol per traversed face   (io.h:109) */
	uint16_t nlists;
	hb_list_streams *lists;
} hb_streams;

/* Streams of a batch of meshes: mesh[i] are the streams of the i-th mesh; their arrays are views into page-locked
 * blocks owned by this object (release everything with hb_batch_streams_free, never hb_streams_free on an entry). */
typedef struct hb_batch_streams {
	uint32_t n;
	hb_streams *mesh;
	void *priv;
} hb_batch_streams;

/* Quantization of one list of every mesh of a batch: what the reference does per mesh with
 * quant::set_bounds (formats/ply/reader.cc:428) + quant::requant (main.cc:108, structs/quant.h:222-242). */
typedef struct hb_quant_req {
	uint32_t list;
	uint8_t new_quant[HB_MAX_COMP]; /* target bits per component, 0 = leave it unquantized */
	uint8_t groups[HB_MAX_COMP];    /* interpretation-group leader of every component (quant.h:54-91) */
} hb_quant_req;
/* requant(clear) of one list of every mesh: bounds = n * 3 rows (min, max, scale of mesh 0, of mesh 1, ...),
 * `stride` bytes each, as read from the .hry headers + set_scale. */
typedef struct hb_dequant_req {
	uint32_t list;
	const void *bounds;
} hb_dequant_req;

typedef struct hb_ctx hb_ctx;

/* ---- context ---------------------------------------------------------------------------- */
int hb_ctx_create(int device, hb_ctx **out);
/* Same, with the context's streams at the device's highest stream priority when high_priority != 0: for the job whose
 * latency matters when two contexts share a GPU (hb_ctx_wait).  Results do not depend on it. */
int hb_ctx_create_prio(int device, int high_priority, hb_ctx **out);
void hb_ctx_destroy(hb_ctx *ctx);
const char *hb_last_error(hb_ctx *ctx); /* ctx may be NULL: last error of hb_ctx_create */
/* Times (ms, CUDA events on the context stream) of the last call: kernels only, and the
 * host<->device copies around them. */
void hb_last_timing(hb_ctx *ctx, float *kernel_ms, float *copy_ms);
/* Row cache (off by default).  While enabled, the device copy of a list's rows survives from one host-buffer call to
 * the next call of the same pipeline -- set_bounds -> requant -> attr_encode, attr_decode -> requant(clear) -- keyed by
 * the host address of the rows (plus row count, stride and quantization state), so the rows cross the link once per
 * direction instead of once per call.  The host rows are still updated by every call.  Contract: between those calls the
 * caller does not modify the rows behind the library's back (the reference's own pipeline, main.cc:98-117, never does).
 * Disabling frees the cached copies. */
int hb_ctx_set_row_cache(hb_ctx *ctx, int enable);
/* Number of kernels this library launched on the context since creation. */
uint64_t hb_kernel_launches(hb_ctx *ctx);
/* Bytes of mesh arrays and attribute rows this library copied host -> device on the context since creation
 * (what an end-to-end measurement counts as its upload). */
uint64_t hb_h2d_bytes(hb_ctx *ctx);

/* Per-kernel timing: while enabled every kernel launch is bracketed by CUDA events on the
 * context stream; the report synchronizes, writes "name launches total_ms\n" lines and resets. */
int hb_ctx_profile(hb_ctx *ctx, int enable);
int hb_ctx_profile_report(hb_ctx *ctx, char *buf, size_t len);
/* Six user events on the context stream to time a region of device-resident stages. */
int hb_ctx_mark(hb_ctx *ctx, int idx);
int hb_ctx_elapsed(hb_ctx *ctx, int from_idx, int to_idx, float *ms);

/* ---- quantization path (host buffers) --------------------------------------------------- */
/* min_row / max_row: `stride` bytes each, laid out like a row of the list with every component
 * in its UNQUANTIZED type (Attr::bounds(), structs/attr.h:33,78-85).  Reproduces the
 * numeric_limits<T>::min() seed of the max row (quant.h:33). */
int hb_bounds(hb_ctx *ctx, const hb_list_desc *list, void *min_row, void *max_row);
/* Converts all rows in place from list->quant[] to new_quant[] (0 = unquantized) and updates
 * list->quant[].  min_row / scale_row as produced by set_bounds / set_scale (quant.h:46-96). */
int hb_requant(hb_ctx *ctx, hb_list_desc *list, const uint8_t *new_quant, const void *min_row,
               const void *scale_row);

/* ---- attribute coder (host buffers) ------------------------------------------------------ */
/* A stream that is all zero is returned as NULL and not copied: reg_vtx / reg_face of a mesh with a single
 * region, type and aux of a list whose emissions are all DATA rows (type_hist[HB_HIST] + type_hist[HB_LHIST] == 0).
 * *out: library-allocated streams in page-locked host memory; hb_streams_free parks the buffers in
 * a process-wide cache for the next call ("Ownership"). */
int hb_attr_encode(hb_ctx *ctx, const hb_mesh_desc *mesh, hb_streams **out);
void hb_streams_free(hb_streams *s);
/* In: lists[l].rows[k] holds the residual of the k-th DATA emission of list l, binding tables
 * filled from the HIST/LHIST symbols.  Out: rows hold the reconstructed attribute values. */
int hb_attr_decode(hb_ctx *ctx, const hb_mesh_desc *mesh);

/* ---- batches of independent meshes (host buffers) -------------------------------------------
 * BASELINE configs[4]; the reference processes one mesh per process (main.cc:93-122).  All meshes of a batch must
 * share one schema (number of lists, their formats, region and binding tables); they differ in size and content.
 * The library cuts the batch into groups, runs every stage ONCE per group over all its meshes (one launch per
 * stage, no host synchronisation inside a stage) and overlaps upload, kernels and download of consecutive groups.
 *
 * hb_encode_batch   per mesh: set_bounds + set_scale + requant of the lists named in q[0..nq), then
 *                   AttrCoder::encode.  bounds_out[k] (may be NULL) receives for request k the rows min, max, scale
 *                   of mesh 0, mesh 1, ... (3 * stride bytes per mesh): the .hry header needs min and max.  The host
 *                   rows are not modified (the quantized rows only exist on the device, as input of the coder).
 * hb_decode_batch   per mesh: AttrDecoder::decode (lists[l].rows in: residual rows; out: values), then
 *                   requant(clear) of the lists named in q[0..nq) -- their rows come back unquantized. */
int hb_encode_batch(hb_ctx *ctx, const hb_mesh_desc *meshes, uint32_t n, const hb_quant_req *q, uint32_t nq,
                    void *const *bounds_out, hb_batch_streams **out);
int hb_decode_batch(hb_ctx *ctx, const hb_mesh_desc *meshes, uint32_t n, const hb_dequant_req *q, uint32_t nq);
void hb_batch_streams_free(hb_batch_streams *b);

/* ---- ingest: twin matching (host buffers) ------------------------------------------------ */
/* Half-edge (f, e) = global index face_off[f] + e runs from its origin to the origin of the next corner
 * of f.  Matches half-edges of opposite direction on the same vertex pair exactly like the reference's
 * Builder does while it reads the faces in file order (first unmatched half-edge of a directed edge
 * waits, a later opposite one merges with it, a later one of the SAME direction stays a border), so
 * non-manifold and degenerate input gives the same table.  Borders are their own twin.
 *   org         origin vertex of every half-edge, one u32 every org_stride bytes: 4 = packed array,
 *               12 = the org field of Conn::edgeorg records (structs/conn.h:73-76); in that case org
 *               may be edges_out itself (in place, what a reader with Builder::automerge = false holds)
 *   edges_out   ne = face_off[nf] records of 12 bytes { u32 org; u32 twin_face; u16 twin_edge; u16 0 },
 *               the layout hb_mesh_desc.edges takes
 * nv = number of vertices (every org must be < nv <= 2^31). */
int hb_twin_match(hb_ctx *ctx, uint32_t nv, uint32_t nf, const uint32_t *face_off, const void *org,
                  uint32_t org_stride, void *edges_out);

/* ---- device-resident pipeline (inputs stay in HBM between stages; used for batches and by
 *      bench.py's kernel-only timing) ---------------------------------------------------- */
typedef struct hb_dmesh hb_dmesh;
int hb_dmesh_upload(hb_ctx *ctx, const hb_mesh_desc *mesh, hb_dmesh **out);
/* n meshes of one schema as ONE device mesh: every stage below then runs once over all of them.  Bounds rows are
 * passed / returned for all meshes back to back (n * stride bytes per row kind). */
int hb_dmesh_upload_batch(hb_ctx *ctx, const hb_mesh_desc *meshes, uint32_t n, hb_dmesh **out);
uint32_t hb_dmesh_segments(hb_dmesh *m); /* number of meshes in the device mesh */
void hb_dmesh_free(hb_dmesh *m);
/* bounds + scale (per `groups`: component j shares the scale of component groups[j], the
 * interpretation-group leader, quant.h:54-91) + requant of list `l`, all on the device. */
int hb_dmesh_quantize(hb_dmesh *m, uint32_t l, const uint8_t *new_quant, const uint8_t *groups);
/* requant(clear): back to the unquantized types using the bounds rows held on the device */
int hb_dmesh_dequantize(hb_dmesh *m, uint32_t l);
/* decode side: min / max rows come from the .hry header, the scale row from set_scale */
int hb_dmesh_set_bounds(hb_dmesh *m, uint32_t l, const void *min_row, const void *max_row, const void *scale_row);
int hb_dmesh_fetch_bounds(hb_dmesh *m, uint32_t l, void *min_row, void *max_row, void *scale_row);
int hb_dmesh_encode(hb_dmesh *m);                    /* kernels only, streams stay on the device */
int hb_dmesh_fetch_streams(hb_dmesh *m, hb_streams **out);              /* single mesh */
int hb_dmesh_fetch_streams_batch(hb_dmesh *m, hb_batch_streams **out);  /* one hb_streams per mesh */
int hb_dmesh_decode(hb_dmesh *m);                    /* kernels only, rows reconstructed in place */
int hb_dmesh_fetch_rows(hb_dmesh *m, uint32_t l, void *rows_out); /* nrows * stride bytes (mesh 0) */
int hb_dmesh_fetch_rows_seg(hb_dmesh *m, uint32_t mesh, uint32_t l, void *rows_out);
/* diagnostics of the speculative vertex reconstruction of list l (8 counters, see hb_api.cu) */
int hb_dmesh_decode_stats(hb_dmesh *m, uint32_t l, uint64_t *out8);
/* keep / restore a device copy of all rows + quantization state (stages work in place) */
int hb_dmesh_snapshot(hb_dmesh *m);
int hb_dmesh_restore(hb_dmesh *m);
int hb_ctx_sync(hb_ctx *ctx);
/* Orders the streams of two contexts of one device: what is queued on `other` so far happens before what is queued on
 * `ctx` from now on.  Two contexts work concurrently on independent meshes (the chain-bound vertex decode of one mesh
 * leaves most SMs idle: the encode of the next mesh fits beside it). */
int hb_ctx_wait(hb_ctx *ctx, hb_ctx *other);

#ifdef __cplusplus
}
#endif
#endif /* HARRY_B200_H */
