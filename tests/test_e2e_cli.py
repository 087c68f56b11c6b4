"""GPU, end to end: the drop-in CLI (harry_b200/host/bin/harry_b200 = the reference's main.cc,
readers, writers, CBM and arithmetic coder compiled unchanged + the GPU attribute path behind the
five swapped calls) must write byte-identical .hry files and byte-identical decoded PLY / OBJ files
to the reference CLI, for every BASELINE config at its small size."""
import filecmp
import os
import subprocess

import pytest

import oracle_lib as ol
from harry_b200 import meshgen

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "harry_b200", "host", "bin", "harry_b200")
needs_bins = pytest.mark.skipif(not (os.path.exists(CLI) and os.path.exists(ol.REF_CLI)),
                                reason="drop-in CLI or reference CLI not built (both need /root/reference at build time)")


def run(exe, *args):
    r = subprocess.run([exe, *args], capture_output=True, text=True)
    assert r.returncode == 0, f"{exe} {' '.join(args)}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}"
    return r.stdout


def gen(workdir, kind):
    if kind == "sphere":
        p = os.path.join(workdir, "e2e_s.ply")
        meshgen.write_ply(p, meshgen.uv_sphere(133, 264))           # config 1: 34 850 vertices
    elif kind == "sphere_noise":
        p = os.path.join(workdir, "e2e_sn.ply")
        meshgen.write_ply(p, meshgen.uv_sphere(60, 97, noise_seed=4))
    elif kind == "poly":
        p = os.path.join(workdir, "e2e_p.ply")
        meshgen.write_ply(p, meshgen.poly_grid(60))                 # config 4: 3 723 vertices
    elif kind == "rgb":
        import cases
        p = os.path.join(workdir, "e2e_rgb.ply")
        meshgen.write_ply(p, cases._rgb_sphere())                   # float xyz + uchar colours (Appendix C.13)
    elif kind == "ints":
        import cases
        p = os.path.join(workdir, "e2e_ints.ply")
        meshgen.write_ply(p, cases._int_irregular())                # int / short / uchar vertex properties, short face property
    elif kind == "obj":
        p = os.path.join(workdir, "e2e_o.obj")
        meshgen.write_obj_latlong(p, 40, 60)                        # config 3: 2 460 vertices
    else:
        p = os.path.join(workdir, "e2e_om.obj")
        meshgen.write_obj_latlong(p, 40, 60, multi_region=True)
    return p


CASES = [
    ("sphere", [], "ply"),                                   # config 1, lossless
    ("sphere", ["-l1", "-q14"], "ply"),                      # config 2 flags at config-1 size
    ("sphere_noise", ["-l1", "-q11"], "ply"),
    ("poly", [], "ply"),                                     # config 4, lossless
    ("poly", ["-l1", "-q12", "-l0", "-q9"], "ply"),
    ("obj", ["-l0", "-q14", "-l2", "-q10"], "obj"),          # config 3
    ("obj", [], "obj"),
    ("obj_multi", ["-l0", "-q14"], "obj"),                   # multi-region variant
    # integer source types: per-component quantization (-a), colours lossless next to quantized coordinates / requantized
    ("rgb", ["-l1", "-a0", "-q12", "-a1", "-q12", "-a2", "-q12"], "ply"),
    ("rgb", ["-l1", "-a0", "-q14", "-a1", "-q14", "-a2", "-q14", "-a3", "-q5", "-a4", "-q5", "-a5", "-q5"], "ply"),
    ("ints", [], "ply"),
    ("ints", ["-l1", "-a0", "-q10", "-a1", "-q10", "-a2", "-q10", "-a3", "-q5", "-a4", "-q13", "-a5", "-q7", "-l0", "-q9"], "ply"),
]


@needs_bins
@pytest.mark.parametrize("kind,flags,ext", CASES, ids=[f"{k}{''.join(f)}" for k, f, _ in CASES])
def test_cli_byte_identical(workdir, kind, flags, ext):
    src = gen(workdir, kind)
    tag = kind + "".join(flags).replace("-", "_")
    ref_hry, our_hry = os.path.join(workdir, tag + "_ref.hry"), os.path.join(workdir, tag + "_b200.hry")
    run(ol.REF_CLI, src, ref_hry, *flags)
    run(CLI, src, our_hry, *flags)
    assert filecmp.cmp(ref_hry, our_hry, shallow=False), ".hry differs from the reference"
    for extra in ([], ["-c"]):
        sfx = "_c" if extra else ""
        ref_out = os.path.join(workdir, f"{tag}_ref{sfx}.{ext}")
        our_out = os.path.join(workdir, f"{tag}_b200{sfx}.{ext}")
        run(ol.REF_CLI, ref_hry, ref_out, *extra)
        run(CLI, our_hry, our_out, *extra)
        assert filecmp.cmp(ref_out, our_out, shallow=False), f"decoded .{ext} differs from the reference ({extra})"


@pytest.mark.skipif(not os.path.exists(CLI), reason="drop-in CLI not built")
def test_cli_error_behaviour(workdir):
    """same error convention as the reference: invalid quantization bits abort the run"""
    src = gen(workdir, "sphere_noise")
    r = subprocess.run([CLI, src, os.path.join(workdir, "x.hry"), "-l1", "-q40"], capture_output=True, text=True)
    assert r.returncode != 0


@pytest.mark.skipif(not os.path.exists(CLI), reason="drop-in CLI not built")
def test_cli_really_runs_the_gpu_path(workdir):
    """no CPU fallback: with an unusable device the drop-in fails in each swapped call instead of
    silently running the reference's CPU code"""
    src = gen(workdir, "sphere_noise")
    env = dict(os.environ, HARRY_B200_DEVICE="99")
    r = subprocess.run([CLI, src, os.path.join(workdir, "y.hry")], capture_output=True, text=True, env=env)
    assert r.returncode != 0 and "harry_b200" in (r.stderr + r.stdout)      # set_bounds in the PLY reader
    # a .hry written by the reference: reading it must hit hb_attr_decode
    if os.path.exists(ol.REF_CLI):
        hry = os.path.join(workdir, "z.hry")
        run(ol.REF_CLI, src, hry)
        r = subprocess.run([CLI, hry, os.path.join(workdir, "z.ply")], capture_output=True, text=True, env=env)
        assert r.returncode != 0 and "harry_b200" in (r.stderr + r.stdout)
