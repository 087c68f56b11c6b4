#!/usr/bin/env python
"""bench.py -- throughput of the attribute path (BASELINE.json metric) on N B200s.

    python bench.py --gpus N --steps K --warmup W            (ours;  N > 1 under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU path)

Workload (config.workload): N = 1 runs BASELINE.json configs[1], one synthetic 10M-vertex UV
sphere (9 999 394 vertices, 19 998 784 triangles), positions quantized `-l1 -q14`, encode + decode.
With N > 1 every rank owns one such mesh (meshes are independent units: sharding by mesh, no
collective, weak scaling).  One step = the whole hot path over the mesh:

    encode side   set_bounds + set_scale + requant(q14)  ->  flatten / ranks / fan gather
                  ->  prediction + residual + byte-plane symbols + histograms
    decode side   fan gather on the decoder's mesh  ->  reconstruction  ->  requant(clear)

`value`   M vertex-attributes/s with all inputs resident in HBM (CUDA events on the library's
          stream around exactly K steps, max over ranks).
`e2e`     the same metric through the host-buffer C ABI (hb_bounds / hb_requant / hb_attr_encode /
          hb_attr_decode) with pinned HOST buffers: H2D of every input and D2H of every output
          inside the timed region.
`roofline` for the kernel with the largest share of the step, from per-launch CUDA events.
`cpu_baseline` the unmodified reference (oracle/_ref, compiled from /root/reference) timed on
          one host core on a bounded sample of the same workload.

Inputs are prepared (untimed) by the reference's own host code -- PLY reader, Cut-Border-Machine
traversal, .hry writer/reader -- because those sequential stages are outside the GPU path.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from harry_b200 import capi, meshgen  # noqa: E402

METRIC = "M vertex-attributes/s encode+decode"
UNIT = "M vertex-attributes/s"
FULL = (2237, 4472)        # configs[1]: 9 999 394 vertices
SAMPLE = (708, 1412)       # bounded CPU sample of the same shape: 998 286 vertices
QBITS = 14


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------
# workload preparation through the reference's host code (untimed)
# ----------------------------------------------------------------------------------------------
class Workload:
    source = "reference host code (oracle/_ref: PLY reader, cbm::encode / cbm::decode, .hry writer/reader)"

    def __init__(self, nr: int, ns: int, workdir: str, keep_ref: bool = False, keep_expected: bool = False):
        import oracle_lib as ol  # input preparation + CPU baseline only
        if not ol.have_ref() or (os.environ.get("HARRY_BENCH_FALLBACK") and not keep_ref):
            if keep_ref:
                raise RuntimeError("oracle/_ref/libharry_ref.so missing")
            self._fallback(nr, ns)
            return
        t0 = time.time()
        self.nr, self.ns = nr, ns
        ply = os.path.join(workdir, f"sphere_{nr}x{ns}.ply")
        meshgen.write_ply(ply, meshgen.uv_sphere(nr, ns))
        rm = ol.RefMesh(ply)
        self.raw = rm.arrays()
        self.loq = [(1, -1, QBITS)]
        self.cpu_times = None
        if keep_ref:
            # time the reference's own functions on this mesh (one core)
            self.cpu_times = rm.time_path(self.loq)
        else:
            rm.requant(self.loq)
            rm.traverse()
        enc = rm.arrays()
        self.new_quant = [la.quants for la in enc.lists]
        self.raw.order, self.raw.order_f, self.raw.edges = enc.order, enc.order_f, enc.edges
        self.enc_expected = rm.attr_encode() if keep_expected else None   # the reference's own AttrCoder on this mesh
        self.hry = os.path.join(workdir, f"sphere_{nr}x{ns}.hry")
        rm.write(self.hry)
        rm.close()
        del enc
        if keep_ref:
            td = ol.ref_time_decode(self.hry)
            self.cpu_times[4], self.cpu_times[5] = td[4], td[5]
        rd = ol.RefMesh(self.hry)
        dec = rd.arrays()
        st = rd.logged_streams()
        self.dec_bounds = []
        for l, la in enumerate(dec.lists):
            rd.set_scale(l)
            self.dec_bounds.append(tuple(rd.bounds_row(l, w, la.stride) for w in (0, 1, 2)))
        rd.close()
        self.dec_expected = [la.rows.copy() for la in dec.lists] if keep_expected else None   # rows the reference decoded
        self.dec = dec.copy()
        self.dec.lists = capi.residual_rows_from_streams(dec, st)
        self.dec.emit_types = [ls.type for ls in st.lists]
        self.n_attrs = self.raw.n_attrs()
        self.nv, self.nf, self.ne = self.raw.nv, self.raw.nf, self.raw.ne
        os.remove(ply)
        log(f"[bench] workload {nr}x{ns}: {self.nv} vertices, {self.nf} faces, {self.n_attrs} attrs, prepared in {time.time() - t0:.1f}s")


    def _fallback(self, nr: int, ns: int):
        """No reference-derived binaries on this box: same mesh, connectivity matched in numpy,
        traversal order = first appearance in the face list (a valid order for the attribute coder,
        not the Cut-Border Machine's), decode input = our own encode output scattered back to rows.
        Labelled in config.workload; parity is not affected (tests use the committed golden vectors)."""
        from harry_b200 import flatten
        t0 = time.time()
        self.nr, self.ns = nr, ns
        Workload.source = "FALLBACK: numpy connectivity + first-appearance order (oracle/_ref absent on this box)"
        self.raw = flatten.mesh_arrays(meshgen.uv_sphere(nr, ns))
        self.loq = [(1, -1, QBITS)]
        self.cpu_times = None
        self.new_quant = [[0] * la.ncomp for la in self.raw.lists]
        self.new_quant[1] = [QBITS] * 3
        ctx = capi.Context(int(os.environ.get("LOCAL_RANK", "0")))
        q = self.raw.copy()
        vl = q.lists[1]
        mn, mx = ctx.bounds(vl)
        sc = float_scale_row(vl, mn, mx)
        ctx.requant(vl, self.new_quant[1], mn, sc)
        st = ctx.attr_encode(q)
        self.dec = q.copy()
        self.dec.lists = capi.residual_rows_encoder_side(q, st)
        self.dec.emit_types = [ls.type for ls in st.lists]
        self.dec_bounds = [(np.zeros(la.stride, np.uint8),) * 3 for la in self.dec.lists]
        self.dec_bounds[1] = (mn, mx, sc)
        ctx.close()
        self.n_attrs = self.raw.n_attrs()
        self.nv, self.nf, self.ne = self.raw.nv, self.raw.nf, self.raw.ne
        log(f"[bench] FALLBACK workload {nr}x{ns}: {self.nv} vertices, prepared in {time.time() - t0:.1f}s")


def float_scale_row(la: capi.ListArrays, mn: np.ndarray, mx: np.ndarray) -> np.ndarray:
    """quant::set_scale (structs/quant.h:46-96) for all-float lists, on the host (a dozen flops)."""
    sc = np.zeros(la.stride, dtype=np.uint8)
    rng = {}
    for j in range(la.ncomp):
        o = la.offsets[j]
        r = np.float32(mx[o:o + 4].view("<f4")[0]) - np.float32(mn[o:o + 4].view("<f4")[0])
        g = la.groups[j]
        rng[g] = max(rng.get(g, np.float32(np.finfo(np.float32).tiny)), np.float32(r))
    for j in range(la.ncomp):
        o = la.offsets[j]
        sc[o:o + 4] = np.array([rng[la.groups[j]]], dtype="<f4").view(np.uint8)
    return sc


# ----------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []   # (arrival time, line)
        self.t0 = None

    def start(self):
        """Started well before the timed region (nvidia-smi needs a few hundred ms to come up, a timed region of
        a few steps is shorter than that); only the samples that arrive inside the region are reported."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def begin_region(self):
        self.t0 = time.perf_counter()

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t1 = time.perf_counter()
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t0 = self.t0 if self.t0 is not None else 0.0
        inside = [ln for (t, ln) in self.lines if t0 <= t <= t1 + 0.02]
        window = "timed region"
        if not inside:   # region shorter than one sampling period: the warm-up steps right before it ran the same work
            inside = [ln for (t, ln) in self.lines[-3:]]
            window = "last samples before the end of the timed region (warm-up steps of the same work)"
        sm, mx, reasons = [], [], set()
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window, "period_ms": 25}


# ----------------------------------------------------------------------------------------------
# algorithmic bytes per launch of each kernel (DESIGN.md "Roofline model")
# ----------------------------------------------------------------------------------------------
def algorithmic_bytes(w: Workload) -> dict:
    nv, ne, A = w.nv, w.ne, w.n_attrs
    s, wd, C, P = 4, 2, 3, 2.0          # source width, storage width, comps/row, parallelograms/vertex
    k5 = A * (wd + wd + 12.0 * P / C)   # own value + residual + triple list amortised
    conn = 16.0 * ne + 4.0 * nv          # every half-edge record once + the rank table
    return {
        "k_bounds_reduce": A * s,
        "k_bounds_reduce_f32<3>": A * s,
        "k_requant": A * (s + wd),
        "k_requant_f32": A * (s + s),          # in place: the 4-byte slot is read and written (the quantized value sits in its low bytes)
        "(k_encode_vtx_packed<T, NC>)": k5,
        "k_vertex_candidates_stage": conn + 4.0 * nv + 12.0 * P * nv,
        "k_vertex_candidates_compact": 2 * 12.0 * P * nv + 4.0 * nv,
        "k_decode_vertex_scan": k5,
        "k_scan_prep": 12.0 * P * nv + 9.0 * nv + 64.0 * nv,
        "k_encode_main<CLS_VTX>": k5,
        "k_decode_vertex_chain": k5,
        "k_decode_vertex_spec3": k5,
        "(k_decode_vertex_spec<T, NC, FP>)": k5,
        "k_flatten_halfedges": 12.0 * ne + 16.0 * ne,
        "k_gather_rp": A * wd * 2,
        "k_scatter_rp": A * wd * 2,
    }


def ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of `kernel` per launch, from the committed summary of the
    last `ncu --set full` capture of this workload (profiles/r01b_*_ncu_full.txt); None if not captured."""
    import glob
    import re
    units = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r01b_*_ncu_full.txt"))):
        cur, rd, wr = None, None, None
        for line in open(path):
            if line.startswith("## "):
                cur, rd, wr = line[3:], None, None
                continue
            m = re.match(r"dram__bytes_(read|write)\.sum = ([0-9.]+) (\w+)", line)
            if m and cur is not None and kernel.split("<")[0].strip("() ") in cur:
                v = float(m.group(2)) * units.get(m.group(3), 1.0)
                if m.group(1) == "read":
                    rd = v
                else:
                    wr = v
                if rd is not None and wr is not None:
                    return rd + wr
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def pinned_like(a: np.ndarray) -> np.ndarray:
    """Copy of `a` in page-locked host memory (torch is only the pinned allocator here)."""
    import torch
    buf = torch.empty(max(1, a.nbytes), dtype=torch.uint8, pin_memory=True)
    out = buf.numpy()[: a.nbytes].view(a.dtype).reshape(a.shape)
    out[...] = a
    pinned_like.keep.append(buf)
    return out


pinned_like.keep = []


def pin_mesh(m: capi.MeshArrays) -> capi.MeshArrays:
    m = m.copy()
    for name in ("edges", "face_off", "order", "order_f", "vtx_regs", "face_regs", "bind_face", "bind_vtx", "bind_corner"):
        a = getattr(m, name)
        if a is not None and a.size:
            setattr(m, name, pinned_like(np.ascontiguousarray(a)))
    for la in m.lists:
        if la.rows.size:
            la.rows = pinned_like(np.ascontiguousarray(la.rows))
    if m.emit_types is not None:   # decode input: the drained type symbols
        m.emit_types = [pinned_like(np.ascontiguousarray(t)) if t is not None and t.size else t for t in m.emit_types]
    return m


def run_ours(args, rank: int, world: int, local_rank: int, dist):
    workdir = tempfile.mkdtemp(prefix="harry_bench_")
    nr, ns = (args.nr, args.ns) if args.nr else FULL
    w = Workload(nr, ns, workdir)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx = capi.Context(local_rank)
    vl = 1
    groups = w.raw.lists[vl].groups
    E = capi.DeviceMesh(ctx, w.raw)
    E.snapshot()
    D = capi.DeviceMesh(ctx, w.dec)
    for l, (mn, mx, sc) in enumerate(w.dec_bounds):
        if w.dec.lists[l].ncomp:
            D.set_bounds(l, mn, mx, sc)
    D.snapshot()
    ctx.sync()

    def step():
        E.restore()
        D.restore()
        ctx.mark(2)
        E.quantize(vl, w.new_quant[vl], groups)
        E.encode()
        ctx.mark(3)
        D.decode()
        D.dequantize(vl)
        ctx.mark(4)

    for _ in range(args.warmup):
        step()
    ctx.sync()
    if dist is not None:
        dist.barrier()
    sampler.begin_region()
    launches0 = ctx.launches()
    ctx.profile(True)
    enc_ms = dec_ms = 0.0
    for _ in range(args.steps):
        step()
        enc_ms += ctx.elapsed(2, 3)
        dec_ms += ctx.elapsed(3, 4)
    ctx.sync()
    prof = ctx.profile_report()
    ctx.profile(False)
    clocks = sampler.stop()
    launches = ctx.launches() - launches0
    total_ms = enc_ms + dec_ms
    if dist is not None:
        import torch
        t = torch.tensor([total_ms, enc_ms, dec_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, enc_ms, dec_ms = (float(x) for x in t.tolist())
        dist.barrier()
    ms_per_step = total_ms / args.steps
    value = world * w.n_attrs / (ms_per_step * 1e-3) / 1e6

    # ---- roofline of the dominant kernel --------------------------------------------------
    alg = algorithmic_bytes(w)
    peak, peak_src = peaks()
    top = max(prof.items(), key=lambda kv: kv[1][1])
    tname, (tn, tms) = top
    per_launch_ms = tms / tn
    bytes_launch = alg.get(tname)
    roof = {"bound": "hbm", "kernel": tname, "launches_per_step": tn / args.steps, "ms_per_launch": per_launch_ms,
            "share_of_step": tms / total_ms if total_ms else None, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
            "traffic": ncu_traffic(tname)}
    if bytes_launch:
        roof["achieved"] = bytes_launch / (per_launch_ms * 1e-3) / 1e9
        roof["frac"] = roof["achieved"] / peak
        roof["algorithmic_bytes_per_launch"] = bytes_launch
    else:
        roof["achieved"] = None
        roof["frac"] = None
    kernels = {}
    for name, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        k = {"launches_per_step": n / args.steps, "ms_per_step": ms / args.steps}
        if alg.get(name):
            k["GBps"] = alg[name] / (ms / n * 1e-3) / 1e9
            k["frac_of_peak"] = k["GBps"] / peak
        kernels[name] = k

    # ---- end to end through the host-buffer C ABI (rank-local, pinned host memory) ----------
    e2e = None
    if not args.no_e2e:
        hraw = pin_mesh(w.raw)
        hdec = pin_mesh(w.dec)
        pristine_raw = [la.rows.copy() for la in w.raw.lists]
        pristine_dec = [la.rows.copy() for la in w.dec.lists]
        n_e2e = max(1, min(args.steps, args.e2e_steps))
        h2d = d2h = 0
        t_e2e = 0.0
        for it in range(n_e2e + 1):
            for la, src in zip(hraw.lists, pristine_raw):
                la.rows[...] = src
                la.quants = [0] * la.ncomp
            for la, src, ref in zip(hdec.lists, pristine_dec, w.dec.lists):
                la.rows[...] = src
                la.quants = list(ref.quants)
            t0 = time.perf_counter()
            la = hraw.lists[vl]
            mn, mx = ctx.bounds(la)
            sc = float_scale_row(la, mn, mx)
            ctx.requant(la, w.new_quant[vl], mn, sc)
            streams, release = ctx.attr_encode_view(hraw)   # the library's page-locked output buffers, as a C++ caller sees them
            ctx.attr_decode(hdec)
            ld = hdec.lists[vl]
            ctx.requant(ld, [0] * ld.ncomp, w.dec_bounds[vl][0], w.dec_bounds[vl][2])
            dt = time.perf_counter() - t0
            if it == 0:      # warm-up
                del streams
                release()
                continue
            t_e2e += dt
            if it == 1:
                rows_b = la.rows.nbytes
                conn_b = sum(getattr(hraw, n).nbytes for n in ("edges", "face_off", "order", "order_f", "vtx_regs", "face_regs", "bind_face", "bind_vtx"))
                # hb_attr_decode of a mesh whose face / corner lists carry no components uploads no face-side arrays (target 1 == HB_VTX)
                vertex_only = all(l.ncomp == 0 or l.nrows == 0 or l.target == 1 for l in hdec.lists)
                dnames = ("edges", "face_off", "order", "vtx_regs", "bind_vtx") if vertex_only else ("edges", "face_off", "order", "vtx_regs", "face_regs", "bind_face", "bind_vtx")
                dconn_b = sum(getattr(hdec, n).nbytes for n in dnames) + \
                    (hdec.order_f.nbytes if hdec.order_f is not None and not vertex_only else 0)
                h2d = rows_b * 2 + conn_b + rows_b + dconn_b + ld.rows.nbytes * 2 + sum(len(t) for t in w.dec.emit_types)
                d2h = rows_b + ld.rows.nbytes * 2 + streams.nbytes_copied   # all-zero streams come back as NULL, not copied
            del streams
            release()
        t_step = t_e2e / n_e2e
        if dist is not None:
            import torch
            t = torch.tensor([t_step], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_step = float(t.item())
        e2e = {"value": world * w.n_attrs / t_step / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": t_step * 1e3, "steps": n_e2e, "timer": "host wall clock around the synchronous C-ABI calls"}

    # ---- CPU baseline: the unmodified reference on a bounded sample, one core ---------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle_lib as ol
        if ol.have_ref():
            cpu = reference_measure(SAMPLE if not args.nr else (args.nr, args.ns), workdir, steps=1)
        else:
            cpu = port_measure(SAMPLE if not args.nr else (args.nr, args.ns))

    E.close()
    D.close()
    twin = None
    if world == 1 and not args.no_twin:
        try:
            twin = run_twin(ctx, w, peak, args.no_cpu)
        except Exception as e:  # the headline line must survive a failure of the extra measurement
            twin = {"error": repr(e)}
    ctx.close()
    batch = None
    if args.batch_meshes > 0 and world == 1:
        import oracle_lib as ol
        if ol.have_ref():
            try:
                batch = run_batch(local_rank, args.batch_meshes, args.batch_threads, 2, workdir)
            except Exception as e:  # the headline line must survive a failure of the extra measurement
                batch = {"error": repr(e)}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u16", "data": "synthetic",
            "config": {"workload": f"configs[1]: UV sphere {nr}x{ns}, {w.nv} vertices / {w.nf} triangles per GPU, float32 xyz, -l1 -q{QBITS}, encode+decode",
                       "inputs_prepared_by": Workload.source,
                       "vertex_attributes_per_gpu": w.n_attrs, "meshes": world, "parallelism": f"mesh-sharded x{world}, no collective",
                       "l2": "inputs (>= 1.3 GB of connectivity + rows per mesh) exceed the 126 MB L2; no explicit flush"},
            "encode_ms_per_step": enc_ms / args.steps, "decode_ms_per_step": dec_ms / args.steps,
            "encode_M_attrs_per_s": world * w.n_attrs / (enc_ms / args.steps * 1e-3) / 1e6 if enc_ms else None,
            "decode_M_attrs_per_s": world * w.n_attrs / (dec_ms / args.steps * 1e-3) / 1e6 if dec_ms else None,
            "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roof, "kernels": kernels,
            "cpu_baseline": cpu, "batch100k": batch, "twin_match": twin,
        }
        print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------
# twin matching of the workload's faces (the step in front of the path)
# ----------------------------------------------------------------------------------------------
def twin_cpu_baseline():
    """Twin matching of the bounded CPU sample on one host core: the reference's own conn::Builder
    (structs/conn.h:172-233, std::unordered_map) driven through oracle/_ref when it is there, else the oracle port."""
    import oracle_lib as ol
    pm = meshgen.uv_sphere(*SAMPLE)
    ne = int(pm.face_idx.shape[0])
    if ol.have_ref_twin():
        _, dt = ol.ref_twin_match(pm.face_off, pm.face_idx)
        kind, what = "reference", "conn::Builder::face_begin / set_org / face_end over all faces (the readers' loop without the parsing)"
    else:
        t0 = time.perf_counter()
        ol.o_twin_match(pm.nv, pm.face_off, pm.face_idx)
        dt = time.perf_counter() - t0
        kind, what = "port", "ho_twin_match (open-addressing restatement of the Builder's unordered_map)"
    return {"value": ne / dt / 1e6, "unit": "M half-edges/s", "cores": 1, "kind": kind,
            "sample": f"UV sphere {SAMPLE[0]}x{SAMPLE[1]} ({ne} half-edges), {what}, single thread"}


def run_twin(ctx, w, peak: float, no_cpu: bool, reps: int = 3):
    """hb_twin_match on the connectivity of the N = 1 workload: host buffers in page-locked memory, wall clock
    around the synchronous call (upload of face_off + origins, three kernels + scan, download of the 12-byte
    records), kernel time from the library's CUDA events, CPU port on the bounded sample beside it."""
    face_off = pinned_like(np.ascontiguousarray(w.raw.face_off))
    org = pinned_like(np.ascontiguousarray(w.raw.edges[:, 0]))
    out = pinned_like(np.zeros((w.ne, 3), np.uint32))
    ctx.twin_match(w.nv, face_off, org, out=out)           # warm-up (memory pool, page-locked registration)
    ctx.profile(True)
    wall = kern = copy = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        ctx.twin_match(w.nv, face_off, org, out=out)
        wall += time.perf_counter() - t0
        k, c = ctx.timing()
        kern += k
        copy += c
    prof = ctx.profile_report()
    ctx.profile(False)
    wall, kern, copy = wall / reps * 1e3, kern / reps, copy / reps
    # algorithmic bytes per half-edge: count reads the origin (4); fill reads it and writes a 16-byte entry (20);
    # resolve reads the origin, its own entry and its partner's, writes the 12-byte record (48)
    alg = {"k_twin_scatter<false>": 4, "k_twin_scatter<true>": 20, "k_twin_resolve<1>": 48}
    kernels = {}
    for name, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        k = {"launches_per_call": n / reps, "ms_per_call": ms / reps}
        if name in alg:
            k["GBps"] = alg[name] * w.ne / (ms / n * 1e-3) / 1e9
            k["frac_of_peak"] = k["GBps"] / peak
        kernels[name] = k
    res = {"half_edges": int(w.ne), "kernel_ms": kern, "copy_ms": copy, "e2e_ms": wall,
           "value": w.ne / (kern * 1e-3) / 1e6 if kern else None, "e2e_value": w.ne / (wall * 1e-3) / 1e6, "unit": "M half-edges/s",
           "h2d_bytes": int(face_off.nbytes + org.nbytes), "d2h_bytes": int(out.nbytes),
           "algorithmic_bytes_per_half_edge": sum(alg.values()), "kernels": kernels,
           "timer": "kernel_ms / copy_ms: CUDA events on the library stream; e2e_ms: host wall clock around the synchronous C-ABI call"}
    del face_off, org, out
    if not no_cpu:
        res["cpu_baseline"] = twin_cpu_baseline()
    return res


# ----------------------------------------------------------------------------------------------
# BASELINE configs[4]: batch of independent 100K-vertex meshes on one GPU (device resident)
# ----------------------------------------------------------------------------------------------
def run_batch(local_rank: int, n_meshes: int, n_threads: int, reps: int, workdir: str):
    """Many independent meshes keep the whole GPU busy even though one mesh's vertex chain only
    occupies one 8-SM cluster: one hb_ctx (= one CUDA stream) per host thread, meshes dealt round
    robin to the threads (harry_b200/shard.py), no exchange of any kind."""
    import concurrent.futures as cf
    from harry_b200 import shard
    distinct = min(n_meshes, 8)
    loads = []
    for k in range(distinct):
        loads.append(BatchMesh(225, 447, 100 + k, workdir))
    plan = shard.shard_plan(n_meshes, n_threads)
    n_attrs = sum(loads[i % distinct].n_attrs for i in range(n_meshes))

    def worker(tid):
        ctx = capi.Context(local_rank)
        mine = []
        for i in plan[tid]:
            bm = loads[i % distinct]
            E = capi.DeviceMesh(ctx, bm.raw)
            E.snapshot()
            D = capi.DeviceMesh(ctx, bm.dec)
            D.set_bounds(1, *bm.dec_bounds)
            D.snapshot()
            mine.append((bm, E, D))
        ctx.sync()
        return ctx, mine

    with cf.ThreadPoolExecutor(n_threads) as pool:
        states = list(pool.map(worker, range(n_threads)))

        def run(st):
            ctx, mine = st
            for bm, E, D in mine:
                E.restore()
                D.restore()
                E.quantize(1, bm.new_quant, bm.groups)
                E.encode()
                D.decode()
                D.dequantize(1)
            ctx.sync()

        list(pool.map(run, states))              # warm-up
        t0 = time.perf_counter()
        for _ in range(reps):
            list(pool.map(run, states))
        dt = (time.perf_counter() - t0) / reps
    launches = sum(st[0].launches() for st in states)
    for ctx, mine in states:
        for _, E, D in mine:
            E.close()
            D.close()
        ctx.close()
    return {"value": n_attrs / dt / 1e6, "unit": UNIT, "meshes": n_meshes, "vertices_per_mesh": loads[0].nv, "host_threads": n_threads,
            "ms_per_batch": dt * 1e3, "ms_per_mesh_amortized": dt * 1e3 / n_meshes, "distinct_meshes": distinct, "gpu_launches_total": int(launches),
            "workload": "configs[4] shape: UV spheres 225x447 (100 130 vertices) with per-mesh seeded radial noise, -l1 -q14, encode+decode, device resident, "
                        "timer: host wall clock around all threads + stream syncs"}


class BatchMesh:
    def __init__(self, nr, ns, seed, workdir):
        import oracle_lib as ol
        ply = os.path.join(workdir, f"b_{seed}.ply")
        meshgen.write_ply(ply, meshgen.uv_sphere(nr, ns, noise_seed=seed))
        rm = ol.RefMesh(ply)
        self.raw = rm.arrays()
        rm.requant([(1, -1, QBITS)])
        rm.traverse()
        enc = rm.arrays()
        self.new_quant = enc.lists[1].quants
        self.groups = self.raw.lists[1].groups
        self.raw.order, self.raw.order_f, self.raw.edges = enc.order, enc.order_f, enc.edges
        hry = ply + ".hry"
        rm.write(hry)
        rm.close()
        rd = ol.RefMesh(hry)
        dec = rd.arrays()
        st = rd.logged_streams()
        rd.set_scale(1)
        self.dec_bounds = tuple(rd.bounds_row(1, w, dec.lists[1].stride) for w in (0, 1, 2))
        rd.close()
        self.dec = dec.copy()
        self.dec.lists = capi.residual_rows_from_streams(dec, st)
        self.dec.emit_types = [ls.type for ls in st.lists]
        self.n_attrs = self.raw.n_attrs()
        self.nv = self.raw.nv
        os.remove(ply)
        os.remove(hry)


# ----------------------------------------------------------------------------------------------
# the reference's CPU path (oracle/_ref = the unmodified reference behind a C harness)
# ----------------------------------------------------------------------------------------------
def reference_measure(shape, workdir, steps=1):
    import oracle_lib as ol
    if not ol.have_ref():
        raise RuntimeError("oracle/_ref/libharry_ref.so missing (built from /root/reference by __graft_entry__.build())")
    nr, ns = shape
    tot = np.zeros(6)
    n_attrs = None
    for _ in range(steps):
        w = Workload(nr, ns, workdir, keep_ref=True)
        tot += np.array(w.cpu_times)
        n_attrs = w.n_attrs
        nv = w.nv
    t = tot / steps
    enc_s = t[0] + t[1] + t[3]
    dec_s = t[4] + t[5]
    return {"value": n_attrs / (enc_s + dec_s) / 1e6, "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": f"UV sphere {nr}x{ns} ({nv} vertices, {n_attrs} attrs), -l1 -q{QBITS}; set_bounds {t[0]*1e3:.0f} ms + requant {t[1]*1e3:.0f} ms + "
                      f"AttrCoder<NullWriter>::encode {t[3]*1e3:.0f} ms (vertices only: {t[2]*1e3:.0f} ms) + AttrDecoder<Replay>::decode {t[4]*1e3:.0f} ms + "
                      f"requant(clear) {t[5]*1e3:.0f} ms; single thread (the reference is single-threaded per mesh)",
            "encode_s": enc_s, "decode_s": dec_s}


def port_measure(shape):
    """CPU baseline when the reference build is absent: the C restatement (oracle/harry_oracle.c),
    one core, on the same bounded sample ("kind": "port")."""
    import oracle_lib as ol
    from harry_b200 import flatten
    nr, ns = shape
    m = flatten.mesh_arrays(meshgen.uv_sphere(nr, ns))
    vl = m.lists[1]
    t0 = time.perf_counter()
    mn, mx = ol.o_bounds(vl)
    sc = ol.o_scale(vl, mn, mx)
    ol.o_requant(vl, [QBITS] * 3, mn, sc)
    st = ol.o_attr_encode(m)
    t1 = time.perf_counter()
    d = m.copy()
    d.lists = capi.residual_rows_encoder_side(m, st)
    t2 = time.perf_counter()
    ol.o_attr_decode(d)
    ol.o_requant(d.lists[1], [0] * 3, mn, sc)
    t3 = time.perf_counter()
    n_attrs = m.n_attrs()
    return {"value": n_attrs / ((t1 - t0) + (t3 - t2)) / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"UV sphere {nr}x{ns} ({m.nv} vertices), -l1 -q{QBITS}, oracle/harry_oracle.c, first-appearance order; encode {1e3*(t1-t0):.0f} ms + decode {1e3*(t3-t2):.0f} ms",
            "encode_s": t1 - t0, "decode_s": t3 - t2}


def _ref_worker(a):
    shape, steps = a
    workdir = tempfile.mkdtemp(prefix="harry_ref_")
    reference_measure(shape, workdir, 1)          # warm-up (page cache, allocator)
    vals, res = [], None
    t0 = time.time()
    for k in range(steps):
        res = reference_measure(shape, workdir, 1)
        vals.append(res["value"])
        if time.time() - t0 > 200:                 # keep the arm within a few minutes
            break
    return float(np.mean(vals)), len(vals), res


def run_reference(args, rank: int, world: int):
    """The reference's own CPU implementation of the path.  It is single-threaded per mesh; with
    N > 1 (N independent meshes, one per GPU in our arm) it gets N host processes, one per mesh."""
    if rank != 0:
        return
    shape = (args.nr, args.ns) if args.nr else SAMPLE
    import oracle_lib as ol
    if not ol.have_ref():
        res = port_measure(shape)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": 1,
                          "warmup": args.warmup, "ms_per_step": 1e3 * (res["encode_s"] + res["decode_s"]), "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "u16", "data": "synthetic",
                          "config": {"workload": "configs[1] shape, bounded sample: " + res["sample"], "parallelism": "1 host thread"},
                          "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                          "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    nproc = max(1, min(world, os.cpu_count() or 1))
    if nproc == 1:
        results = [_ref_worker((shape, args.steps))]
    else:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(nproc) as pool:
            results = pool.map(_ref_worker, [(shape, args.steps)] * nproc)
    v = float(sum(r[0] for r in results)) * (world / nproc)
    steps = min(r[1] for r in results)
    res = results[0][2]
    enc_dec_s = res["encode_s"] + res["decode_s"]
    sample = res["sample"] + (f"; {nproc} processes, one mesh each" if nproc > 1 else "")
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
        "ms_per_step": enc_dec_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": {"workload": f"configs[1] shape, bounded sample: {sample}", "parallelism": f"{nproc} host process(es), one mesh each (the reference is single-threaded per mesh)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": nproc, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nr", type=int, default=0, help="override the sphere size (rings); default = configs[1]")
    ap.add_argument("--ns", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-twin", action="store_true", help="skip the extra measurement of hb_twin_match (SURVEY 8f row f2)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--batch-meshes", type=int, default=64, help="extra measurement: batch of independent 100K-vertex meshes (0 = skip)")
    ap.add_argument("--batch-threads", type=int, default=8)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        dist = dist_mod
    run_ours(args, rank, world, local_rank, dist)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
