#!/usr/bin/env python
"""bench.py -- throughput of the attribute path (BASELINE.json metric) on N B200s.

    python bench.py --gpus N --steps K --warmup W            (ours;  N > 1 under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU path)

Workload (config.workload): BASELINE.json configs[1], one synthetic 10M-vertex UV sphere
(9 999 394 vertices, 19 998 784 triangles), positions quantized `-l1 -q14`, encode + decode.
With N > 1 every rank owns one such mesh (meshes are independent units: sharding by mesh, no
collective, weak scaling).  EVERY run (N = 1, 2, 4, 8) also measures BASELINE configs[4], the mesh
batch (`batch` object: 1250 independent 100K-vertex meshes per GPU through the batched entry points,
`value`, `e2e` and -- on rank 0 at N = 1 -- the reference on all host cores).  One step = the whole
hot path over the mesh:

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
          one host core on the SAME mesh (one step of the full workload, taken while the inputs are prepared).
`cli`     wall clock and the reference's own phase prints of `oracle/_ref/harry` and the drop-in
          `harry_b200/host/bin/harry_b200` on configs[1] (encode, decode -c), N = 1 only.
`configs` device-resident / e2e lines for configs[0], [2], [3] with their dominant kernel, N = 1 only.

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
        self.ply = ply
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
        log(f"[bench] workload {nr}x{ns}: {self.nv} vertices, {self.nf} faces, {self.n_attrs} attrs, prepared in {time.time() - t0:.1f}s")

    def cpu_baseline(self):
        """The reference's own functions on THIS mesh, one core, one step (timed during the preparation)."""
        t = self.cpu_times
        enc_s = t[0] + t[1] + t[3]
        dec_s = t[4] + t[5]
        return {"value": self.n_attrs / (enc_s + dec_s) / 1e6, "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": f"one step of the full workload: UV sphere {self.nr}x{self.ns} ({self.nv} vertices, {self.n_attrs} attrs), -l1 -q{QBITS}; set_bounds {t[0]*1e3:.0f} ms + "
                          f"requant {t[1]*1e3:.0f} ms + AttrCoder<NullWriter>::encode {t[3]*1e3:.0f} ms (vertices only: {t[2]*1e3:.0f} ms) + AttrDecoder<Replay>::decode {t[4]*1e3:.0f} ms + "
                          f"requant(clear) {t[5]*1e3:.0f} ms; single thread (the reference is single-threaded per mesh)",
                "encode_s": enc_s, "decode_s": dec_s}


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
        "k_requant_f32<1>": A * (s + wd),      # s + w per attribute (the in-place kernel moves s + s: read-modify-write of the 4-byte slot)
        "k_requant_f32<2>": A * (s + wd),
        "k_requant_f32<0>": A * (s + wd),
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
    last `ncu --set full` capture of this workload (profiles/r02_kernels_10m_ncu_full.txt); None if not captured."""
    import glob
    import re
    units = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_kernels_10m_ncu_full.txt"))):
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


def allreduce_max(dist, local_rank, vals):
    if dist is None:
        return list(vals)
    import torch
    t = torch.tensor(list(vals), dtype=torch.float64, device=f"cuda:{local_rank}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def arm_config(nr, ns, n_attrs, world):
    """The `config` object of the bench line: the SAME dict in the GPU arm and in the `--impl reference` arm (both run this
    workload: one configs[1] mesh per GPU / per reference process, prepared once, K timed steps)."""
    nv, nf = (nr - 1) * ns + 2, 2 * ns * (nr - 1)
    return {"workload": f"configs[1]: UV sphere {nr}x{ns}, {nv} vertices / {nf} triangles per mesh, float32 xyz, -l1 -q{QBITS}, encode+decode",
            "vertex_attributes_per_gpu": int(n_attrs), "meshes": int(world),
            "parallelism": f"{world} mesh(es), one per GPU (reference arm: one per host process), sharded by mesh, no collective",
            "l2": "inputs (>= 1.3 GB of connectivity + rows per mesh) exceed the 126 MB L2; no explicit flush"}


def link_probe(dist, local_rank, world, nbytes=1 << 30, reps=3):
    """Host -> device bandwidth of one GPU while the others idle, and of every GPU when all ranks copy at once: the
    end-to-end lines are link-bound, so their scaling over N is the platform's (GPUs sharing a PCIe uplink), measured here
    with plain torch copies of 1 GiB from page-locked memory (CUDA events)."""
    import torch
    dev = torch.device(f"cuda:{local_rank}")
    host = dst = None
    try:
        host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    except Exception:
        pass
    ok = torch.tensor([1.0 if dst is not None else 0.0], dtype=torch.float64, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)      # every rank or none: nobody waits at a barrier for a rank that gave up
    if ok.item() < 1.0:
        return {"error": "could not allocate the probe buffers on every rank"}

    def timed():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dst.copy_(host, non_blocking=True)   # warm-up
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            dst.copy_(host, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        return reps * nbytes / (a.elapsed_time(b) * 1e-3) / 1e9

    alone = []
    for r in range(world):               # one rank at a time
        dist.barrier()
        alone.append(timed() if r == dist.get_rank() else 0.0)
    dist.barrier()
    together = timed()
    t = torch.tensor(alone + [together, -together], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    v = t.tolist()
    alone_all, tmax, tmin = v[:world], v[world], -v[world + 1]
    s = torch.tensor([together], dtype=torch.float64, device=dev)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    del host, dst
    return {"h2d_GBps_alone_per_gpu": [round(x, 1) for x in alone_all], "h2d_GBps_concurrent_min": round(tmin, 1), "h2d_GBps_concurrent_max": round(tmax, 1),
            "h2d_GBps_concurrent_sum": round(float(s.item()), 1),
            "note": "plain 1 GiB page-locked copies, no library code: what the platform gives N ranks at once; the e2e lines cannot scale past it"}


def kernel_table(prof, alg, peak, steps, total_ms):
    kernels = {}
    for name, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        k = {"launches_per_step": n / steps, "ms_per_step": ms / steps, "share_of_step": ms / total_ms if total_ms else None}
        if alg.get(name):
            k["GBps"] = alg[name] / (ms / n * 1e-3) / 1e9
            k["frac_of_peak"] = k["GBps"] / peak
        kernels[name] = k
    return kernels


def run_ours(args, rank: int, world: int, local_rank: int, dist):
    workdir = tempfile.mkdtemp(prefix="harry_bench_")
    nr, ns = (args.nr, args.ns) if args.nr else FULL
    import oracle_lib as ol
    want_cpu = rank == 0 and world == 1 and not args.no_cpu and ol.have_ref()
    w = Workload(nr, ns, workdir, keep_ref=want_cpu)
    sampler = ClockSampler(local_rank)
    sampler.start()
    # two contexts (= two streams) on the GPU: the encode of the mesh and the decode of the decoder-side mesh are
    # independent jobs, and the chain-bound vertex decode of ONE mesh occupies 48 of the 148 SMs -- the encode fits beside it
    ctx = capi.Context(local_rank)
    ctx_d = ctx if args.one_stream else capi.Context(local_rank, high_priority=not args.flat_priority)   # the chain-bound job goes first
    vl = 1
    groups = w.raw.lists[vl].groups
    E = capi.DeviceMesh(ctx, w.raw)
    E.snapshot()
    D = capi.DeviceMesh(ctx_d, w.dec)
    for l, (mn, mx, sc) in enumerate(w.dec_bounds):
        if w.dec.lists[l].ncomp:
            D.set_bounds(l, mn, mx, sc)
    D.snapshot()
    ctx.sync()
    ctx_d.sync()

    def step():
        E.restore()
        D.restore()
        if ctx_d is not ctx:            # both streams start the step together, behind both restores
            ctx.wait(ctx_d)
            ctx_d.wait(ctx)
        ctx.mark(2)
        ctx_d.mark(0)
        E.quantize(vl, w.new_quant[vl], groups)
        E.encode()
        ctx.mark(3)
        D.decode()
        D.dequantize(vl)
        ctx_d.mark(1)
        if ctx_d is not ctx:
            ctx.wait(ctx_d)
        ctx.mark(4)                     # both jobs done

    for _ in range(args.warmup):
        step()
    ctx.sync()
    ctx_d.sync()
    if dist is not None:
        dist.barrier()
    sampler.begin_region()
    launches0 = ctx.launches() + (ctx_d.launches() if ctx_d is not ctx else 0)
    enc_ms = dec_ms = total_ms = 0.0
    for _ in range(args.steps):
        step()
        enc_ms += ctx.elapsed(2, 3)
        dec_ms += ctx_d.elapsed(0, 1)
        total_ms += ctx.elapsed(2, 4)
    ctx.sync()
    ctx_d.sync()
    clocks = sampler.stop()
    launches = ctx.launches() + (ctx_d.launches() if ctx_d is not ctx else 0) - launches0
    # per-kernel durations (roofline of the dominant kernel): CUDA events around every launch, in extra steps that run
    # the two jobs one after the other -- a kernel timed while the other stream runs beside it measures the neighbour too
    ctx.profile(True)
    ctx_d.profile(True)
    prof_steps = 3
    for _ in range(prof_steps):
        E.restore()
        D.restore()
        E.quantize(vl, w.new_quant[vl], groups)
        E.encode()
        ctx_d.wait(ctx)
        D.decode()
        D.dequantize(vl)
        ctx.wait(ctx_d)
    ctx.sync()
    ctx_d.sync()
    prof = ctx.profile_report()
    if ctx_d is not ctx:
        for k, (n, ms) in ctx_d.profile_report().items():
            n0, ms0 = prof.get(k, (0, 0.0))
            prof[k] = (n0 + n, ms0 + ms)
    ctx.profile(False)
    ctx_d.profile(False)
    serial_ms = sum(ms for _, ms in prof.values()) / prof_steps
    total_ms, enc_ms, dec_ms = allreduce_max(dist, local_rank, [total_ms, enc_ms, dec_ms])
    if dist is not None:
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
    roof = {"bound": "hbm", "kernel": tname, "launches_per_step": tn / prof_steps, "ms_per_launch": per_launch_ms,
            "share_of_step": (tms / prof_steps) / serial_ms if serial_ms else None, "share_of": "sum of the kernel times of a step",
            "peak": peak, "peak_source": peak_src, "unit": "GB/s", "traffic": ncu_traffic(tname)}
    if bytes_launch:
        roof["achieved"] = bytes_launch / (per_launch_ms * 1e-3) / 1e9
        roof["frac"] = roof["achieved"] / peak
        roof["algorithmic_bytes_per_launch"] = bytes_launch
    else:
        roof["achieved"] = None
        roof["frac"] = None
    if tname == "k_decode_vertex_scan":
        roof["note"] = ("latency-bound dependency chain of ONE mesh (SURVEY 8d: reported as such, not hidden); the same kernel over a batch of "
                        "meshes is in batch.kernels")
    kernels = kernel_table(prof, alg, peak, prof_steps, serial_ms * prof_steps)

    # ---- end to end through the host-buffer C ABI (rank-local, pinned host memory) ----------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, ctx, w, vl, dist, local_rank, world)

    E.close()
    D.close()
    cpu = w.cpu_baseline() if want_cpu else None
    twin = None
    if world == 1 and not args.no_twin:
        try:
            twin = run_twin(ctx, w, peak, args.no_cpu)
        except Exception as e:  # the headline line must survive a failure of the extra measurement
            twin = {"error": repr(e)}
    n_attrs, nv, nf = w.n_attrs, w.nv, w.nf
    ply, hry = w.ply, w.hry
    del w, E, D
    pinned_like.keep.clear()

    # ---- BASELINE configs[4]: the mesh batch, on every rank at every N ------------------------------
    batch = None
    if ctx_d is not ctx and not args.flat_priority:
        # the batch keeps every SM busy from both streams: no job to favour there
        ctx_d.close()
        ctx_d = capi.Context(local_rank)
    if args.batch_meshes > 0 and ol.have_ref():
        try:
            batch = run_batch(args, ctx, ctx_d, rank, world, local_rank, dist, workdir, peak)
        except Exception as e:  # the headline line must survive a failure of the extra measurement
            import traceback
            log(traceback.format_exc())
            batch = {"error": repr(e)}
    pinned_like.keep.clear()
    configs = None
    if rank == 0 and world == 1 and not args.no_configs and ol.have_ref():
        try:
            configs = run_other_configs(ctx, workdir, peak)
        except Exception as e:
            configs = {"error": repr(e)}
    if ctx_d is not ctx:
        ctx_d.close()
    ctx.close()
    link = None
    if world > 1:
        try:
            link = link_probe(dist, local_rank, world)
        except Exception as e:
            link = {"error": repr(e)}
    cli = None
    if rank == 0 and world == 1 and not args.no_cli:
        try:
            cli = run_cli(ply, workdir, n_attrs)
        except Exception as e:
            cli = {"error": repr(e)}
    for f in (ply, hry):
        try:
            os.remove(f)
        except OSError:
            pass

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u16", "data": "synthetic",
            "config": arm_config(nr, ns, n_attrs, world),
            "notes": {"inputs_prepared_by": Workload.source,
                      "streams": ("one stream" if args.one_stream else "two streams of the same GPU: quantize + encode of the mesh on one, decode + dequantize of the decoder-side mesh "
                                  "on the other, which has the higher stream priority (independent inputs; the chain-bound vertex decode of one mesh occupies 48 of 148 SMs, "
                                  "the encode job takes what it leaves); ms_per_step = both done; "
                                  "encode_ms / decode_ms = each job on its own stream while the other runs"),
                      "batch_config": "configs[4] (the mesh batch north_star scales on) is measured in the same run at every N: see `batch`"},
            "encode_ms_per_step": enc_ms / args.steps, "decode_ms_per_step": dec_ms / args.steps,
            "encode_M_attrs_per_s": world * n_attrs / (enc_ms / args.steps * 1e-3) / 1e6 if enc_ms else None,
            "decode_M_attrs_per_s": world * n_attrs / (dec_ms / args.steps * 1e-3) / 1e6 if dec_ms else None,
            "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roof, "kernels": kernels,
            "cpu_baseline": cpu, "batch": batch, "cli": cli, "configs": configs, "twin_match": twin,
        }
        if link is not None:
            out["link"] = link
        print(json.dumps(out), flush=True)


def run_e2e(args, ctx, w, vl, dist, local_rank, world):
    """configs[1] through the calls the drop-in adapter makes (hb_bounds, hb_requant, hb_attr_encode, hb_attr_decode,
    hb_requant(clear)) on page-locked HOST buffers: every upload and download inside the timer, all `--steps` steps."""
    hraw = pin_mesh(w.raw)
    hdec = pin_mesh(w.dec)
    pristine_raw = [la.rows.copy() for la in w.raw.lists]
    pristine_dec = [la.rows.copy() for la in w.dec.lists]
    n_e2e = max(1, args.steps)
    h2d = d2h = 0
    t_e2e = 0.0
    CALLS = ("hb_bounds", "hb_requant", "hb_attr_encode", "hb_attr_decode", "hb_requant(clear)")
    per_call = {k: 0.0 for k in CALLS}
    per_step = []
    ctx.set_row_cache(True)      # what the CLI adapter does (host/bridge.cc): rows stay on the device between the calls of a pipeline
    n_warm = max(1, args.warmup)   # untimed steps first: the second or third call can still grow the stream-ordered pool
    for it in range(n_e2e + n_warm):
        for la, src in zip(hraw.lists, pristine_raw):
            la.rows[...] = src
            la.quants = [0] * la.ncomp
        for la, src, ref in zip(hdec.lists, pristine_dec, w.dec.lists):
            la.rows[...] = src
            la.quants = list(ref.quants)
        up0 = ctx.h2d_bytes()
        t0 = time.perf_counter()
        la = hraw.lists[vl]
        mn, mx = ctx.bounds(la)
        t1 = time.perf_counter()
        sc = float_scale_row(la, mn, mx)
        ctx.requant(la, w.new_quant[vl], mn, sc)
        t2 = time.perf_counter()
        streams, release = ctx.attr_encode_view(hraw)   # the library's page-locked output buffers, as a C++ caller sees them
        t3 = time.perf_counter()
        ctx.attr_decode(hdec)
        t4 = time.perf_counter()
        ld = hdec.lists[vl]
        ctx.requant(ld, [0] * ld.ncomp, w.dec_bounds[vl][0], w.dec_bounds[vl][2])
        t5 = time.perf_counter()
        dt = t5 - t0
        up1 = ctx.h2d_bytes()
        if it < n_warm:      # warm-up
            del streams
            release()
            continue
        t_e2e += dt
        per_step.append([round(x * 1e3, 2) for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)])
        for k, d in zip(CALLS, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
            per_call[k] += d * 1e3 / n_e2e
        if it == n_warm:
            rows_b = la.rows.nbytes
            # counted by the library at its host -> device copies: rows once for set_bounds (requant and encode find them on the device),
            # once for decode; the connectivity of the encoder and of the decoder mesh (12 bytes per half-edge each); single-region
            # arrays are cleared on the device; the decoder mesh goes without face-side arrays, the encoder mesh without the face
            # order when the device finds that no face row is bound twice
            h2d = up1 - up0
            d2h = rows_b + ld.rows.nbytes * 2 + streams.nbytes_copied   # all-zero streams come back as NULL, not copied
        del streams
        release()
    ctx.set_row_cache(False)
    t_step = allreduce_max(dist, local_rank, [t_e2e / n_e2e])[0]
    return {"value": world * w.n_attrs / t_step / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": t_step * 1e3, "ms_per_call": {k: round(v, 3) for k, v in per_call.items()}, "ms_per_call_by_step": per_step,
            "steps": n_e2e, "timer": "host wall clock around the synchronous C-ABI calls",
            "calls": "hb_bounds, hb_requant, hb_attr_encode, hb_attr_decode, hb_requant(clear) -- what the drop-in CLI adapter calls"}


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
# BASELINE configs[4]: batch of independent 100K-vertex meshes (sharded by mesh: every rank its own batch)
# ----------------------------------------------------------------------------------------------
BATCH_SHAPE = (225, 447)   # 100 130 vertices, 200 256 triangles


class BatchMesh:
    """One mesh of the batch, prepared (untimed) by the reference's host code like the N = 1 workload."""

    def __init__(self, nr, ns, seed, workdir, keep_ref=False):
        import oracle_lib as ol
        ply = os.path.join(workdir, f"b_{os.getpid()}_{seed}.ply")
        meshgen.write_ply(ply, meshgen.uv_sphere(nr, ns, noise_seed=seed))
        rm = ol.RefMesh(ply)
        self.raw = rm.arrays()
        self.cpu_times = None
        if keep_ref:
            rm.snapshot()
            rm.time_encode_step([(1, -1, QBITS)])      # warm-up (also runs the traversal once)
            rm.restore()
            self.cpu_times = rm.time_encode_step([(1, -1, QBITS)])
        else:
            rm.requant([(1, -1, QBITS)])
            rm.traverse()
        enc = rm.arrays()
        self.new_quant = enc.lists[1].quants
        self.groups = self.raw.lists[1].groups
        self.raw.order, self.raw.order_f, self.raw.edges = enc.order, enc.order_f, enc.edges
        hry = ply + ".hry"
        rm.write(hry)
        rm.close()
        if keep_ref:
            rdec = ol.RefDecoder(hry)
            rdec.step()
            td = rdec.step()
            rdec.close()
            self.cpu_times[4], self.cpu_times[5] = td[4], td[5]
        rd = ol.RefMesh(hry)
        dec = rd.arrays()
        st = rd.logged_streams()
        rd.set_scale(1)
        self.dec_bounds = tuple(rd.bounds_row(1, w, dec.lists[1].stride) for w in (0, 1, 2))
        rd.requant([], clear=True)
        self.deq_rows = rd.arrays().lists[1].rows.copy()    # what the reference's `-c` decode yields
        rd.close()
        self.dec = dec.copy()
        self.dec.lists = capi.residual_rows_from_streams(dec, st)
        self.dec.emit_types = [ls.type for ls in st.lists]
        self.n_attrs = self.raw.n_attrs()
        self.nv = self.raw.nv
        os.remove(ply)
        os.remove(hry)


def _batch_cpu_worker(a):
    seed, workdir = a
    bm = BatchMesh(*BATCH_SHAPE, seed, workdir, keep_ref=True)
    t = bm.cpu_times
    return bm.n_attrs, t[0] + t[1] + t[3] + t[4] + t[5]


def batch_cpu_baseline(workdir):
    """The reference on ALL host cores: one process per mesh (main.cc:93-122 is one mesh per process), disjoint meshes."""
    import multiprocessing as mp
    nproc = max(1, os.cpu_count() or 1)
    nproc = min(nproc, 64)
    with mp.get_context("spawn").Pool(nproc) as pool:
        res = pool.map(_batch_cpu_worker, [(9000 + k, workdir) for k in range(nproc)])
    rate = sum(n / t for n, t in res) / 1e6
    return {"value": rate, "unit": UNIT, "cores": nproc, "kind": "reference",
            "sample": f"{nproc} processes in parallel, one 100 130-vertex mesh each (disjoint seeds), one timed step after a warm-up step: set_bounds + requant + "
                      f"AttrCoder<NullWriter>::encode + AttrDecoder<Replay>::decode + requant(clear); mean {1e3 * float(np.mean([t for _, t in res])):.0f} ms per mesh and core"}


def run_batch(args, ctx, ctx_d, rank, world, local_rank, dist, workdir, peak):
    """configs[4] on this rank: `--batch-meshes` independent meshes (cycled from `--batch-distinct` prepared ones).
    value: device resident -- `--batch-resident` meshes live in HBM as device meshes of one group each (one launch per
    stage over the group); a step runs quantize + encode + decode + dequantize over every group, `passes` times.
    e2e: hb_encode_batch + hb_decode_batch on page-locked host buffers, all meshes, uploads and downloads inside."""
    import ctypes as C
    n_meshes, distinct = args.batch_meshes, min(args.batch_distinct, args.batch_meshes)
    t0 = time.time()
    loads = [BatchMesh(*BATCH_SHAPE, 1000 * rank + 100 + k, workdir) for k in range(distinct)]
    log(f"[bench] batch: {distinct} distinct meshes prepared in {time.time() - t0:.1f}s")
    attrs_per_mesh = loads[0].n_attrs
    ne_mesh = loads[0].raw.ne
    group = max(1, min(n_meshes, int(args.batch_group_half_edges // max(1, ne_mesh))))
    resident = min(n_meshes, max(group, args.batch_resident // group * group))
    n_groups = resident // group
    passes = max(1, round(n_meshes / resident))
    n_value = passes * n_groups * group           # meshes per device-resident step
    # ---- device resident ---------------------------------------------------------------------
    Es, Ds = [], []
    for g in range(n_groups):
        idx = [(g * group + k) % distinct for k in range(group)]
        E = capi.DeviceMesh(ctx, [loads[i].raw for i in idx])
        E.snapshot()
        D = capi.DeviceMesh(ctx_d, [loads[i].dec for i in idx])
        D.set_bounds(1, *(np.stack([loads[i].dec_bounds[wh] for i in idx]) for wh in (0, 1, 2)))
        D.snapshot()
        Es.append(E)
        Ds.append(D)
    ctx.sync()
    ctx_d.sync()
    bm0 = loads[0]

    def step():
        # encode groups on one stream, decode groups on the other (independent meshes)
        if ctx_d is not ctx:
            ctx.wait(ctx_d)
            ctx_d.wait(ctx)
        ctx.mark(2)
        for _ in range(passes):
            for E, D in zip(Es, Ds):
                E.restore()
                D.restore()
                E.quantize(1, bm0.new_quant, bm0.groups)
                E.encode()
                D.decode()
                D.dequantize(1)
        if ctx_d is not ctx:
            ctx.wait(ctx_d)
        ctx.mark(3)

    for _ in range(args.warmup):
        step()
    ctx.sync()
    ctx_d.sync()
    # parity spot check inside the bench: first and last mesh of the last group against the reference's decode
    idx_last = [((n_groups - 1) * group + k) % distinct for k in range(group)]
    for seg in (0, group - 1):
        if not np.array_equal(Ds[-1].fetch_rows(1, seg), loads[idx_last[seg]].deq_rows):
            raise RuntimeError("batch decode differs from the reference")
    if dist is not None:
        dist.barrier()
    launches0 = ctx.launches() + (ctx_d.launches() if ctx_d is not ctx else 0)
    ms = 0.0
    for _ in range(args.steps):
        step()
        ms += ctx.elapsed(2, 3)
    ctx.sync()
    ctx_d.sync()
    launches = ctx.launches() + (ctx_d.launches() if ctx_d is not ctx else 0) - launches0
    # per-kernel durations: one extra step with the two jobs one after the other (see run_ours)
    ctx.profile(True)
    ctx_d.profile(True)
    for _ in range(passes):
        for E, D in zip(Es, Ds):
            E.restore()
            E.quantize(1, bm0.new_quant, bm0.groups)
            E.encode()
            ctx_d.wait(ctx)
            D.restore()
            D.decode()
            D.dequantize(1)
            ctx.wait(ctx_d)
    ctx.sync()
    ctx_d.sync()
    prof = ctx.profile_report()
    if ctx_d is not ctx:
        for k, (n, kms) in ctx_d.profile_report().items():
            n0, ms0 = prof.get(k, (0, 0.0))
            prof[k] = (n0 + n, ms0 + kms)
    ctx.profile(False)
    ctx_d.profile(False)
    serial_ms = sum(kms for _, kms in prof.values())
    ms = allreduce_max(dist, local_rank, [ms])[0]
    if dist is not None:
        dist.barrier()
    ms_step = ms / args.steps
    value = world * n_value * attrs_per_mesh / (ms_step * 1e-3) / 1e6
    A = n_value * attrs_per_mesh
    nvv, nee = n_value * bm0.nv, n_value * ne_mesh
    alg = {"k_bounds_reduce_f32<3>": A * 4, "k_requant_f32<1>": A * 6, "k_requant_f32<2>": A * 6, "(k_encode_vtx_packed<T, NC>)": A * 12.0,
           "k_decode_vertex_scan": A * 12.0, "k_flatten_halfedges": 28.0 * nee * 2, "k_vertex_candidates_stage": (16.0 * nee + 8.0 * nvv + 24.0 * nvv) * 2,
           "k_vertex_candidates_compact": (48.0 * nvv + 4.0 * nvv) * 2, "k_scan_prep": (24.0 + 9.0 + 64.0) * nvv}
    # per-launch figures: a kernel is launched passes * n_groups (* 2 for stages shared by encode and decode) times per step
    kernels = {}
    for name, (n, kms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        k = {"launches_per_step": n, "ms_per_step": kms, "share_of_kernel_time": kms / serial_ms if serial_ms else None}
        if alg.get(name):
            k["GBps"] = alg[name] / (kms * 1e-3) / 1e9
            k["frac_of_peak"] = k["GBps"] / peak
        kernels[name] = k
    for E, D in zip(Es, Ds):
        E.close()
        D.close()
    del Es, Ds
    out = {"value": value, "unit": UNIT, "ms_per_step": ms_step, "meshes_per_gpu_per_step": n_value, "vertices_per_mesh": bm0.nv, "attrs_per_mesh": attrs_per_mesh,
           "meshes_per_launch": group, "resident_meshes": resident, "passes_per_step": passes, "distinct_meshes": distinct, "n_gpus": world,
           "gpu_launches_per_step": launches / args.steps, "launches_per_mesh": launches / args.steps / n_value,
           "ms_per_mesh_amortized": ms_step / n_value, "kernels": kernels, "kernel_ms_serial_step": serial_ms,
           "streams": "encode groups on one stream, decode groups on a second stream of the same GPU (independent meshes); kernels{} from one extra step "
                      "that runs them one after the other",
           "workload": f"configs[4]: UV spheres {BATCH_SHAPE[0]}x{BATCH_SHAPE[1]} (100 130 vertices) with per-mesh seeded radial noise, -l1 -q{QBITS}, encode+decode; "
                       f"{n_value} meshes per GPU and step, sharded by mesh over {world} GPU(s), no collective",
           "timer": "CUDA events on the library stream around each step, max over ranks"}
    # ---- end to end: host buffers, pipelined groups ----------------------------------------------
    if not args.no_e2e:
        lib = ctx.lib
        hraw = [pin_mesh(bm.raw) for bm in loads]
        hdec = [pin_mesh(bm.dec) for bm in loads]
        pristine = [bm.dec.lists[1].rows for bm in loads]
        n_e2e = n_meshes
        # decode works in place on the rows of every mesh: each of the n meshes gets its own page-locked row buffer
        dec_meshes = []
        for i in range(n_e2e):
            m = hdec[i % distinct].copy()            # shares the connectivity arrays, owns its rows
            m.lists[1].rows = pinned_like(pristine[i % distinct])
            dec_meshes.append(m)
        enc_descs = capi._desc_array([hraw[i % distinct] for i in range(n_e2e)])
        dec_descs = capi._desc_array(dec_meshes)
        req = (capi.QuantReq * 1)()
        req[0].list = 1
        for j, v in enumerate(bm0.new_quant):
            req[0].new_quant[j] = v
        for j, v in enumerate(bm0.groups):
            req[0].groups[j] = v
        stride = bm0.raw.lists[1].stride
        bounds_out = pinned_like(np.zeros((n_e2e, 3, stride), np.uint8))
        bptr = (C.c_void_p * 1)(bounds_out.ctypes.data)
        dq_bounds = pinned_like(np.stack([np.stack(loads[i % distinct].dec_bounds) for i in range(n_e2e)]))
        dreq = (capi.DequantReq * 1)()
        dreq[0].list = 1
        dreq[0].bounds = dq_bounds.ctypes.data
        t_sum, d2h_streams = 0.0, 0
        n_steps = max(1, min(args.steps, 5))
        n_warm = max(1, min(args.warmup, 3))
        for it in range(n_steps + n_warm):
            for i, m in enumerate(dec_meshes):       # fresh residual rows (untimed)
                m.lists[1].rows[...] = pristine[i % distinct]
            up0 = ctx.h2d_bytes()
            t0 = time.perf_counter()
            bp = C.POINTER(capi.BatchStreams)()
            ctx._check(lib.hb_encode_batch(ctx.h, enc_descs, n_e2e, req, 1, bptr, C.byref(bp)), "hb_encode_batch")
            ctx._check(lib.hb_decode_batch(ctx.h, dec_descs, n_e2e, dreq, 1), "hb_decode_batch")
            dt = time.perf_counter() - t0
            if it == n_warm:
                h2d_step = ctx.h2d_bytes() - up0 + n_e2e * 3 * stride   # + the bounds rows of the dequantization requests
                st = capi.streams_struct_to_py(bp.contents.mesh[n_e2e - 1], copy=False)
                d2h_streams = st.nbytes_copied
                if not np.array_equal(dec_meshes[n_e2e - 1].lists[1].rows, loads[(n_e2e - 1) % distinct].deq_rows):
                    raise RuntimeError("hb_decode_batch differs from the reference")
                del st
            lib.hb_batch_streams_free(bp)
            if it >= n_warm:
                t_sum += dt
        t_step = allreduce_max(dist, local_rank, [t_sum / n_steps])[0]
        d0 = hdec[0]
        out["e2e"] = {"value": world * n_e2e * attrs_per_mesh / t_step / 1e6, "unit": UNIT, "ms_per_step": t_step * 1e3, "meshes_per_gpu_per_step": n_e2e, "steps": n_steps,
                      "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": int(n_e2e * (d2h_streams + 3 * stride + d0.lists[1].rows.nbytes)),
                      "calls": "hb_encode_batch (set_bounds + set_scale + requant + encode) + hb_decode_batch (decode + requant(clear)), page-locked host buffers, "
                               "groups of meshes pipelined over copy / compute / download streams",
                      "timer": "host wall clock around the two synchronous C-ABI calls, max over ranks"}
        del dec_meshes, hraw, hdec, bounds_out, dq_bounds
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = batch_cpu_baseline(workdir)
    return out


# ----------------------------------------------------------------------------------------------
# the other BASELINE configs, one line each (N = 1): lossless float vertex lists, OBJ corner lists, polygons
# ----------------------------------------------------------------------------------------------
def run_other_configs(ctx, workdir, peak, reps=5):
    import cases

    def obj(d):
        pth = os.path.join(d, "cfg3.obj")
        meshgen.write_obj_latlong(pth, 400, 600)
        return pth

    todo = [
        ("configs[0]: UV sphere 133x264 (34 850 vertices), lossless float32", (lambda d: cases._ply(d, "cfg1.ply", meshgen.uv_sphere(133, 264))), []),
        ("configs[2]: OBJ lat-long sphere 401x600 (240 600 vertices, 480 000 triangles) with vt + vn corner lists, -l0 -q14 -l2 -q10", obj, [(0, -1, 14), (2, -1, 10)]),
        ("configs[3]: polygon grid n=600 (tri / quad / 5- / 6-gons + non-manifold fin), per-vertex and per-face floats, lossless", (lambda d: cases._ply(d, "cfg4.ply", meshgen.poly_grid(600))), []),
        ("mixed source types (SURVEY App. C.13): UV sphere 400x800 (319 202 vertices), float xyz at -a0..2 -q14 next to lossless uchar red / green / blue in ONE list",
         (lambda d: cases._ply(d, "cfg_rgb.ply", meshgen.with_typed_props(meshgen.uv_sphere(400, 800, noise_seed=8), seed=3, vtx=(("red", np.uint8), ("green", np.uint8), ("blue", np.uint8))))),
         [(1, 0, 14), (1, 1, 14), (1, 2, 14)]),
    ]
    out = []
    for title, gen, loq in todo:
        c = cases.Case(workdir, "cfg_" + str(len(out)), gen, loq)
        n_attrs = c.raw.n_attrs()
        raw = c.raw.copy()
        raw.order, raw.order_f, raw.edges = c.enc.order, c.enc.order_f, c.enc.edges
        E = capi.DeviceMesh(ctx, raw)
        E.snapshot()
        din = c.decode_input()
        D = capi.DeviceMesh(ctx, din)
        for l, (mn, mx) in enumerate(c.dec_bounds):
            if din.lists[l].ncomp:
                D.set_bounds(l, mn, mx, c.deq_scale[l] if c.deq is not None else None)
        D.snapshot()
        qreq = [(l, c.enc.lists[l].quants, la.groups) for l, la in enumerate(raw.lists) if la.ncomp and c.enc.lists[l].quants != la.quants]

        def step():
            E.restore()
            D.restore()
            ctx.mark(2)
            for l, nq, gr in qreq:
                E.quantize(l, nq, gr)
            E.encode()
            ctx.mark(3)
            D.decode()
            for l, _, _ in qreq:
                D.dequantize(l)
            ctx.mark(4)

        for _ in range(3):
            step()
        ok, why = E.fetch_streams().equal(c.enc_streams)
        want = c.deq if c.deq is not None else c.dec
        for l, la in enumerate(want.lists):
            ok = ok and (la.ncomp == 0 or np.array_equal(D.fetch_rows(l), la.rows))
        ctx.profile(True)
        enc = dec = 0.0
        for _ in range(reps):
            step()
            enc += ctx.elapsed(2, 3)
            dec += ctx.elapsed(3, 4)
        prof = ctx.profile_report()
        ctx.profile(False)
        top = max(prof.items(), key=lambda kv: kv[1][1])
        dstats = [D.decode_stats(l) for l, la in enumerate(din.lists) if la.ncomp and la.target == 1]
        E.close()
        D.close()
        # e2e: the adapter's calls on host buffers (pageable here: what the CLI passes)
        t0 = time.perf_counter()
        hr = raw.copy()
        for l, nq, gr in qreq:
            la = hr.lists[l]
            mn, mx = ctx.bounds(la)
            ctx.requant(la, nq, mn, c.raw_scale[l])
        ctx.attr_encode(hr)
        hd = c.decode_input()
        ctx.attr_decode(hd)
        for l, _, _ in qreq:
            ctx.requant(hd.lists[l], [0] * hd.lists[l].ncomp, c.dec_bounds[l][0], c.deq_scale[l])
        t_e2e = time.perf_counter() - t0
        out.append({"workload": title, "vertex_attributes": n_attrs, "parity_vs_reference": bool(ok),
                    "value": n_attrs / ((enc + dec) / reps * 1e-3) / 1e6, "unit": UNIT, "encode_ms": enc / reps, "decode_ms": dec / reps,
                    "e2e_value": n_attrs / t_e2e / 1e6, "e2e_ms": t_e2e * 1e3,
                    "vertex_decode_stats": dstats,
                    "dominant_kernel": {"name": top[0], "launches_per_step": top[1][0] / reps, "ms_per_step": top[1][1] / reps, "share_of_step": top[1][1] / (enc + dec)}})
    return out


# ----------------------------------------------------------------------------------------------
# the CLIs side by side on configs[1] (main.cc:98-120 phase prints)
# ----------------------------------------------------------------------------------------------
def run_cli(ply, workdir, n_attrs):
    import re
    ref = os.path.join(ROOT, "oracle", "_ref", "harry")
    ours = os.path.join(ROOT, "harry_b200", "host", "bin", "harry_b200")
    if not (os.path.exists(ref) and os.path.exists(ours) and os.path.exists(ply)):
        return {"unavailable": "CLI binaries (built from /root/reference in the build container) or the input are missing"}

    def run(binary, a, b, flags):
        t0 = time.perf_counter()
        r = subprocess.run([binary, a, b] + flags, capture_output=True, text=True)
        dt = time.perf_counter() - t0
        if r.returncode != 0:
            raise RuntimeError(f"{binary} failed: {r.stderr[-300:]}")
        phases = {}
        for line in r.stdout.replace("\r", "\n").splitlines():
            m = re.match(r"\s*(Reading input|Quantization|Writing output) took:?\s*([0-9.]+)", line)
            if m:
                phases[m.group(1)] = float(m.group(2))
        return dt, phases

    out = {"workload": "configs[1]: in.ply -> out.hry -l1 -q14, then out.hry -> back.ply -c", "unit": "s wall clock per process"}
    files = {}
    for name, binary in (("reference", ref), ("harry_b200", ours)):
        hry = os.path.join(workdir, f"cli_{name}.hry")
        back = os.path.join(workdir, f"cli_{name}.ply")
        te, pe = run(binary, ply, hry, ["-l1", "-q14"])
        td, pd = run(binary, hry, back, ["-c"])
        out[name] = {"encode_s": te, "decode_s": td, "encode_phases": pe, "decode_phases": pd,
                     "M_vertex_attributes_per_s": n_attrs / (te + td) / 1e6}
        files[name] = (hry, back)
    same = all(open(files["reference"][k], "rb").read() == open(files["harry_b200"][k], "rb").read() for k in (0, 1))
    out["byte_identical_outputs"] = bool(same)
    out["speedup_wall_clock"] = (out["reference"]["encode_s"] + out["reference"]["decode_s"]) / (out["harry_b200"]["encode_s"] + out["harry_b200"]["decode_s"])
    for pair in files.values():
        for f in pair:
            os.remove(f)
    return out


# ----------------------------------------------------------------------------------------------
# the reference's CPU path (oracle/_ref = the unmodified reference behind a C harness)
# ----------------------------------------------------------------------------------------------
def port_measure(shape):
    """CPU baseline when the reference build is absent: the C restatement (oracle/harry_oracle.c),
    one core, on a bounded sample ("kind": "port")."""
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
    """One process = one mesh of configs[1] (the reference is single-threaded per mesh): prepare ONCE, then time
    `steps` steps of the path on that mesh (rows and formats restored between steps, untimed)."""
    shape, steps, warmup, budget_s, seed = a
    import oracle_lib as ol
    workdir = tempfile.mkdtemp(prefix="harry_ref_")
    nr, ns = shape
    ply = os.path.join(workdir, f"ref_{os.getpid()}.ply")
    meshgen.write_ply(ply, meshgen.uv_sphere(nr, ns))
    rm = ol.RefMesh(ply)
    os.remove(ply)
    n_attrs, nv = rm.arrays().n_attrs(), 0
    loq = [(1, -1, QBITS)]
    rm.snapshot()
    rm.time_encode_step(loq)                 # first pass: traversal (untimed part) + the quantized state the file needs
    hry = os.path.join(workdir, "ref.hry")
    rm.write(hry)
    dec = ol.RefDecoder(hry)                 # (the file stays: every step reads its header again)
    times, t_start = [], time.time()
    for k in range(warmup + steps):
        rm.restore()
        t = rm.time_encode_step(loq)
        td = dec.step()
        t[4], t[5] = td[4], td[5]
        if k >= warmup:
            times.append(t)
        if time.time() - t_start > budget_s and len(times) >= 1:
            break
    dec.close()
    rm.close()
    os.remove(hry)
    tm = np.mean(np.array(times), axis=0)
    return n_attrs, [float(x) for x in tm], len(times)


def run_reference(args, rank: int, world: int):
    """The reference's own CPU implementation of the path on the SAME config as our arm: configs[1], one mesh per GPU of
    our arm = one host process per mesh here (the reference is single-threaded per mesh, main.cc:93-122)."""
    if rank != 0:
        return
    shape = (args.nr, args.ns) if args.nr else FULL
    import oracle_lib as ol
    if not ol.have_ref():
        res = port_measure(SAMPLE)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": 1,
                          "warmup": args.warmup, "ms_per_step": 1e3 * (res["encode_s"] + res["decode_s"]), "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "u16", "data": "synthetic",
                          "config": {"workload": "configs[1] shape, bounded sample (oracle/_ref absent): " + res["sample"], "parallelism": "1 host thread"},
                          "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                          "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    nproc = max(1, min(world, os.cpu_count() or 1))
    warm = min(args.warmup, 1)
    job = (shape, args.steps, warm, 150.0, 0)
    if nproc == 1:
        results = [_ref_worker(job)]
    else:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(nproc) as pool:
            results = pool.map(_ref_worker, [job] * nproc)
    n_attrs = results[0][0]
    rates = [r[0] / (r[1][0] + r[1][1] + r[1][3] + r[1][4] + r[1][5]) / 1e6 for r in results]
    v = float(sum(rates)) * (world / nproc)
    steps = min(r[2] for r in results)
    t = results[0][1]
    step_s = t[0] + t[1] + t[3] + t[4] + t[5]
    sample = (f"the full workload: UV sphere {shape[0]}x{shape[1]} ({n_attrs} attrs per mesh), -l1 -q{QBITS}, prepared once, {steps} timed step(s) per mesh; set_bounds {t[0]*1e3:.0f} ms + "
              f"requant {t[1]*1e3:.0f} ms + AttrCoder<NullWriter>::encode {t[3]*1e3:.0f} ms + AttrDecoder<Replay>::decode {t[4]*1e3:.0f} ms + requant(clear) {t[5]*1e3:.0f} ms"
              + (f"; {nproc} processes, one mesh each" if nproc > 1 else "; single thread (the reference is single-threaded per mesh)"))
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": arm_config(shape[0], shape[1], n_attrs, world),
        "notes": {"processes": f"{nproc} host process(es), one mesh each (the reference is single-threaded per mesh)"},
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
    ap.add_argument("--one-stream", action="store_true", help="encode and decode one after the other on one stream (default: two streams, independent meshes)")
    ap.add_argument("--flat-priority", action="store_true", help="A/B: the decode context at the same stream priority as the encode context")
    ap.add_argument("--no-cli", action="store_true", help="skip the side-by-side run of the two CLIs on configs[1]")
    ap.add_argument("--no-configs", action="store_true", help="skip the lines for configs[0], [2], [3]")
    ap.add_argument("--batch-meshes", type=int, default=1250, help="configs[4]: independent 100K-vertex meshes per GPU and step (0 = skip)")
    ap.add_argument("--batch-distinct", type=int, default=16, help="distinct prepared meshes the batch cycles through")
    ap.add_argument("--batch-resident", type=int, default=800, help="meshes kept resident in HBM for the device-resident batch value")
    ap.add_argument("--batch-group-half-edges", type=int, default=224 << 20,
                    help="half-edges per device mesh of the device-resident batch value (group of meshes run by one launch per stage; 391 spheres of 100K vertices = "
                         "1173 chains = eight per SM in the scan decoder).  The host-buffer calls cut their own groups (112M half-edges: shorter pipeline fill and drain)")
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
        # one rank per GPU: stay on the CPUs of the GPU's NUMA node, so that the page-locked buffers allocated from here on
        # (first touch) sit behind the same PCIe root as the GPU (no-op on a flat topology)
        from harry_b200 import shard
        shard.bind_rank_to_gpu_node(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        dist = dist_mod
    run_ours(args, rank, world, local_rank, dist)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
