"""Parity tests proper: the CUDA library on a B200, called through the C ABI by the package's
Engine / drop-in functions, against the CPU oracle and the reference-generated golden fixtures.
Same checks as the emulation suite (tests/parity_checks.py) plus full-size and end-to-end cases.
"""
import argparse
import contextlib
import io

import numpy as np
import pytest

from oracle import numpy_port as P
from spatial_light_modulator_module_b200 import host_logic as hl, synthetic
from tests import parity_checks as pc

pytestmark = pytest.mark.gpu


def make_engine(shape, precision, max_batch):
    from spatial_light_modulator_module_b200.engine import Engine
    return Engine(shape, precision, max_batch)


def ns(**kw):
    base = dict(incomming_intensity="uniform", tolerance=0, max_loops=10, gif=False, gif_skip=1, gif_type="i",
                print_info=False, plot_error=False, initial_guess="random", random_seed=42, white_attention=1,
                learning_rate=0.005, unsettle=0, correspond_to2pi=256)
    base.update(kw)
    return argparse.Namespace(**base)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def test_native_library_is_the_compute_path():
    import ctypes
    from spatial_light_modulator_module_b200 import _ffi
    lib = _ffi.load()
    assert isinstance(lib, ctypes.CDLL) and lib.slm_version() >= 100
    eng = make_engine((128, 128), "fp32", 1)
    n0 = eng.launch_count()
    eng.gs(synthetic.noise_target((128, 128)), 3)
    # setup(2), row, max pre-pass, 3 col, 2 row, final row, intensity (a context for ONE plane: plain column kernels, 1 launch)
    assert eng.launch_count() - n0 == 2 + 1 + 1 + 3 + 2 + 1 + 1
    eng.close()
    eng = make_engine((1024, 1024), "fp32", 2)                    # the tile pipeline: intensity = transform kept + scaling
    n0 = eng.launch_count()
    eng.gs(synthetic.noise_target((1024, 1024)), 3)
    assert eng.launch_count() - n0 == 2 + 1 + 1 + 3 + 2 + 1 + 2
    eng.close()


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("shape", [(64, 64), (128, 192), (192, 256), (512, 512), (768, 1024), (1024, 1024),
                                   (2048, 2048), (64, 4096), (4096, 64)])
def test_fft2_matches_scipy(shape, precision):
    pc.check_fft2(make_engine, shape, precision)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("kind", ["noise", "shapes"])
@pytest.mark.parametrize("shape", [(128, 128), (192, 256), (768, 1024)])
def test_gs_teacher_forced(shape, kind, precision):
    pc.check_gs_teacher_forced(make_engine, shape, precision, kind, steps=(0, 1, 4) if shape[0] < 768 else (0, 2))


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("kind", ["noise", "shapes", "traps"])
def test_gs_device_setup(kind, precision):
    pc.check_gs_device_setup(make_engine, (128, 128), precision, kind)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("name", ["gd_noise_random_128x128", "gd_shapes_fourier_192x256", "gd_traps_unsettle_128x128",
                                  "gd_noise_wa2int_128x128", "gd_noise_wa05_128x128", "gd_shapes_old_128x128",
                                  "gd_shapes_unnormed_128x128", "gd_shapes_zeros_128x128", "gd_shapes_ones_128x128",
                                  "gd_traps_tol_128x128"])
def test_gd_matches_reference_golden(golden, name, precision):
    pc.check_gd_golden(make_engine, golden, name, precision)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_gs_tolerance_and_batch(precision):
    pc.check_gs_tolerance_and_batch(make_engine, precision)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_gd_tolerance_and_batch(precision):
    pc.check_gd_tolerance_and_batch(make_engine, precision)


@pytest.mark.parametrize("shape", [(1024, 128), (768, 1024), (1024, 1024)])
def test_warp_column_kernel_tolerance_and_batch(shape):
    """Planes that stop early inside a batch, on the warp-per-column kernel (GS) and its fused GD pass."""
    pc.check_gs_tolerance_and_batch(make_engine, "fp32", shape=shape)
    pc.check_gd_tolerance_and_batch(make_engine, "fp32", shape=shape, loops=6)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_gs_real_valued_targets(golden, precision):
    pc.check_gs_real_targets(make_engine, golden, precision)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_illumination(precision):
    pc.check_illumination(make_engine, precision)


def test_analytic_and_quantisers_bit_exact(golden):
    pc.check_analytic_and_quantisers(make_engine, golden)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_expected_outcome(golden, precision):
    pc.check_expected_outcome(make_engine, golden, precision)


# ---- full-size cases against the reference's recorded curves (tests/golden/*_curves.npz) -------------
def test_gd_full_size_curve_fp64(golden):
    """Config 2 shape (768x1024, 100 iterations) free-running against the unmodified reference."""
    g = golden("gd_noise_768x1024_curves")
    t = synthetic.noise_target((768, 1024), seed=0)
    holo, exp, errs = quiet(__import__("spatial_light_modulator_module_b200.algorithms", fromlist=["x"]).gradient_descent,
                            t, ns(max_loops=100, precision="fp64"))
    assert len(errs) == 100
    assert np.max(np.abs(np.array(errs) - g["errors"]) / g["errors"]) < 1e-9
    sub = (slice(None, None, 16), slice(None, None, 16))
    assert pc.circ(holo[sub], g["hologram_sub"]).max() < 1e-8
    assert np.abs(exp[sub] - g["expected_sub"]).max() < 1e-9 * g["expected_sub"].max()


def test_gd_full_size_curve_fp32(golden):
    g = golden("gd_noise_768x1024_curves")
    t = synthetic.noise_target((768, 1024), seed=0)
    from spatial_light_modulator_module_b200 import algorithms
    holo, exp, errs = quiet(algorithms.gradient_descent, t, ns(max_loops=100, precision="fp32"))
    assert np.max(np.abs(np.array(errs) - g["errors"]) / g["errors"]) < 1e-3
    sub = (slice(None, None, 16), slice(None, None, 16))
    d = pc.circ(holo[sub], g["hologram_sub"])
    assert np.mean(d < 1e-3) >= 0.999                      # north star: 1e-3 rad
    assert np.abs(exp[sub] - g["expected_sub"]).max() < 1e-3 * g["expected_sub"].max()   # north star: 1e-3


@pytest.mark.parametrize("shape", [(1024, 1024), (768, 1024), (1024, 512)])
def test_gd_fused_fourier_pass_vs_oracle(shape):
    """fp32 columns of 1024 / 768 points: the warp-per-column kernel, GD with the plane max taken inside ONE
    Fourier-plane pass (every tile of a plane in flight at once) -- against the oracle, in a batch."""
    pc.check_gd_vs_oracle(make_engine, shape, "fp32", "noise", loops=5, batch=3)


def test_gd_fused_equals_two_pass():
    """The fused pass and the two-pass form (max pass that keeps the transform + gradient pass) run the same
    arithmetic: identical curves and holograms.  The two-pass library instance lives in a child process."""
    import subprocess, sys, os, tempfile
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from spatial_light_modulator_module_b200.engine import Engine\n"
        "from spatial_light_modulator_module_b200 import synthetic, host_logic as hl\n"
        "shape=(1024,1024); eng=Engine(shape,'fp32',2)\n"
        "t=np.stack([synthetic.noise_target(shape,seed=i) for i in range(2)])\n"
        "x0=np.stack([hl.host_initial_guess('random',shape,42)]*2)\n"
        "during,_=hl.learning_rate_schedule(0.005,0,6)\n"
        "r,_=eng.gd(t,x0,during,6)\n"
        "np.savez(sys.argv[1], e=np.array(r.errors), h=eng.to_host(r.hologram), x=eng.to_host(r.expected))\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, env in (("fused", {}), ("two_pass", {"SLM_NO_FUSED_GD": "1"})):
            path = os.path.join(tmp, name + ".npz")
            subprocess.run([sys.executable, "-c", code, path], check=True, env={**os.environ, **env}, timeout=600)
            out[name] = dict(np.load(path))
    for k in ("e", "h", "x"):
        np.testing.assert_array_equal(out["fused"][k], out["two_pass"][k])


def test_gs_traps_full_size_from_reference_phasor(golden):
    """Trap movie frame (768x1024, 50 iterations): free-running from the reference's first phasor."""
    g = golden("gs_traps_768x1024_curves")
    pts = [tuple(p) for p in g["trap_points"]]
    t = synthetic.traps_target((768, 1024), pts)
    st = P.gs_setup(t)
    B0 = P.gs_first_phasor(st)
    for precision, tol in (("fp64", 1e-6), ("fp32", 1e-3)):
        eng = make_engine((768, 1024), precision, 1)
        res = eng.gs(t, 50, phasor0=B0)
        e = res.errors[0]
        assert len(e) == 50
        rel = np.abs(e - g["errors"]) / g["errors"]
        # iteration 1 reads the phase of A = ifft2(D) on the node lines of the two-trap field, where the
        # reference's value is pocketfft rounding noise: its error differs by a few %
        assert rel[1] < 0.05
        assert np.max(np.delete(rel, 1)) < tol
        eng.close()


def test_gs_512_config1_statistics(golden):
    """Config 1 (512x512, 20 iterations, dense noise): the trajectory is chaotic, so the
    free-running device run is compared with the reference curve at the first iteration tightly
    and at the end statistically."""
    g = golden("gs_noise_512x512_curves")
    from spatial_light_modulator_module_b200 import algorithms
    t = synthetic.noise_target((512, 512), seed=0)
    for precision in ("fp32", "fp64"):
        holo, exp, errs = quiet(algorithms.gerchberg_saxton, t, ns(max_loops=20, precision=precision))
        assert len(errs) == 20 and holo.shape == (512, 512) and holo.dtype == np.float64
        assert abs(errs[0] - g["errors"][0]) < 1e-5 * g["errors"][0]
        assert 0.5 * g["errors"][-1] < errs[-1] < 1.5 * g["errors"][-1] and errs[-1] < 0.5 * errs[0]
        assert all(isinstance(e, np.float64) for e in errs)


# ---- size-independent properties at the bench sizes -------------------------------------------------
@pytest.mark.parametrize("shape", [(1024, 1024), (2048, 2048)])
def test_fft_roundtrip_and_parseval(shape):
    eng = make_engine(shape, "fp32", 1)
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    X = eng.fft2(x)
    xr = eng.to_host(eng.fft2(X, inverse=True))
    assert np.abs(xr - x).max() < 5e-6 * np.abs(x).max()
    Xh = eng.to_host(X).astype(np.complex128)
    assert abs((np.abs(Xh) ** 2).sum() / (shape[0] * shape[1]) - (np.abs(x.astype(np.complex128)) ** 2).sum()) \
        < 1e-5 * (np.abs(x) ** 2).sum()
    eng.close()


def test_gs_invariants_1024():
    """At 1024^2: |hologram| <= pi, max(expected) == max(target), the returned error equals
    error_f(expected, target) recomputed on the host, and the error decreases on a trap target."""
    from spatial_light_modulator_module_b200 import algorithms
    t = synthetic.traps_target((1024, 1024), [(300, 200), (700, 500), (512, 900)])
    holo, exp, errs = quiet(algorithms.gerchberg_saxton, t, ns(max_loops=15, precision="fp32"))
    assert np.all(np.abs(holo) <= np.pi)
    assert abs(exp.max() - 255.0) < 1e-9
    assert abs(algorithms.error_f(exp, t, t.size) - errs[-1]) < 1e-6 * errs[-1] + 1e-12
    assert abs(errs[-1] - errs[-2]) < 1e-6 * errs[-1]          # converged to a fixed point
    # the hologram really produces the expected outcome: |fft2(exp(i*h))|^2 peaks at the traps
    from spatial_light_modulator_module_b200 import generate_hologram as gh
    prev = gh.expected_outcome(holo, 255, precision="fp32")
    top = np.argsort(prev.ravel())[-3:]
    assert set(map(tuple, np.argwhere(t == 255))) == set(zip(*np.unravel_index(top, t.shape)))


def test_graph_replay_gives_the_same_results():
    """Loops are replayed from CUDA graphs: a small run is launched plainly the first time, recorded the second time its
    arguments are seen and replayed from the kept graph afterwards; a large batch records 20 iterations per graph.  Same
    bits and the same launch count every time, and the same as with SLM_NO_GRAPH=1 (child process)."""
    import os, subprocess, sys, tempfile
    import torch
    shape = (1024, 1024)
    t = synthetic.noise_target(shape, seed=4)
    x0 = hl.host_initial_guess("random", shape, 42).astype(np.complex64)
    during, _ = hl.learning_rate_schedule(0.005, 0, 12)
    eng = make_engine(shape, "fp32", 8)
    td = torch.from_numpy(t[None]).cuda()
    x0d = torch.from_numpy(x0[None]).cuda()
    x = torch.empty_like(x0d)
    runs = []
    for rep in range(4):                       # same device buffers every time: plain, recorded, replayed, replayed
        x.copy_(x0d)
        n0 = eng.launch_count()
        res, _ = eng.gd(td, x, during, 12, norms=np.array([float(t.max())]))
        runs.append((res.errors[0].copy(), eng.to_host(res.hologram).copy(), eng.launch_count() - n0))
        g = eng.gs(td, 12, norms=np.array([float(t.max())]))
        runs[-1] += (g.errors[0].copy(), eng.to_host(g.hologram).copy())
        del res, g
    for r in runs[1:]:
        np.testing.assert_array_equal(r[0], runs[0][0])
        np.testing.assert_array_equal(r[1], runs[0][1])
        assert r[2] == runs[0][2]
        np.testing.assert_array_equal(r[3], runs[0][3])
        np.testing.assert_array_equal(r[4], runs[0][4])
    # a batch (graphs of 20 iterations + a remainder launched plainly) against the same batch without graphs
    tb = np.stack([synthetic.noise_target(shape, seed=10 + i) for i in range(8)])
    during50, _ = hl.learning_rate_schedule(0.005, 0, 50)
    xb = np.stack([x0] * 8)
    res, _ = eng.gd(tb, xb, during50, 50)
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from spatial_light_modulator_module_b200.engine import Engine\n"
        "from spatial_light_modulator_module_b200 import synthetic, host_logic as hl\n"
        "shape=(1024,1024); eng=Engine(shape,'fp32',8)\n"
        "tb=np.stack([synthetic.noise_target(shape,seed=10+i) for i in range(8)])\n"
        "x0=hl.host_initial_guess('random',shape,42).astype(np.complex64)\n"
        "during,_=hl.learning_rate_schedule(0.005,0,50)\n"
        "r,_=eng.gd(tb,np.stack([x0]*8),during,50)\n"
        "np.savez(sys.argv[1], e=np.array(r.errors), h=eng.to_host(r.hologram))\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "plain.npz")
        subprocess.run([sys.executable, "-c", code, path], check=True, env={**os.environ, "SLM_NO_GRAPH": "1"}, timeout=600)
        plain = dict(np.load(path))
    np.testing.assert_array_equal(np.array(res.errors), plain["e"])
    np.testing.assert_array_equal(eng.to_host(res.hologram), plain["h"])
    eng.close()


def test_batched_equals_single_1024():
    eng = make_engine((768, 1024), "fp32", 4)
    frames = synthetic.movie_frames(4)
    res = eng.gs(frames, 8)
    hb = eng.to_host(res.hologram)
    for k in (0, 3):
        r1 = eng.gs(frames[k], 8)
        np.testing.assert_array_equal(eng.to_host(r1.hologram)[0], hb[k])
        np.testing.assert_array_equal(r1.errors[0], res.errors[k])
    eng.close()


# ---- drop-in surface ----------------------------------------------------------------------------------
def test_dropin_error_behaviour():
    from spatial_light_modulator_module_b200 import algorithms
    t = synthetic.traps_target((128, 128))
    with pytest.raises(UnboundLocalError):
        quiet(algorithms.gerchberg_saxton, t, ns(max_loops=0))
    with pytest.raises(UnboundLocalError):
        quiet(algorithms.gradient_descent, t, ns(max_loops=0))
    with pytest.raises(ValueError, match="unknown type of initial guess"):
        quiet(algorithms.gradient_descent, t, ns(initial_guess="nope"))
    with pytest.raises(AttributeError):
        a = ns()
        del a.random_seed
        quiet(algorithms.gradient_descent, t, a)
    with pytest.raises(ZeroDivisionError):
        quiet(algorithms.gradient_descent, t, ns(max_loops=1, unsettle=5))
    with pytest.raises(ValueError):
        quiet(algorithms.gerchberg_saxton, np.zeros((100, 100), np.uint8), ns())   # unsupported plane shape
    # all-zero target: one iteration, error 0.0, zero hologram (algorithms.py:29 stops on `0 > 0`)
    holo, exp, errs = quiet(algorithms.gerchberg_saxton, np.zeros((128, 128), np.uint8), ns(max_loops=7))
    assert errs == [0.0] and not holo.any() and not exp.any()


def test_dropin_unsettle_mutates_args(golden):
    from spatial_light_modulator_module_b200 import algorithms
    g = golden("gd_traps_unsettle_128x128")
    a = ns(max_loops=12, unsettle=2, learning_rate=0.01, precision="fp64")
    holo, exp, errs = quiet(algorithms.gradient_descent, g["target"], a)
    assert a.learning_rate == float(g["final_learning_rate"])
    assert np.max(np.abs(np.array(errs) - g["errors"]) / g["errors"]) < 1e-9


def test_dropin_progress_output(capsys):
    from spatial_light_modulator_module_b200 import algorithms
    algorithms.gerchberg_saxton(synthetic.traps_target((128, 128)), ns(max_loops=4, print_info=True))
    out = capsys.readouterr().out
    assert "\rloop 4/4" in out and "number of loops: 4" in out and "error: " in out


def test_dropin_analytic_and_display(golden, tmp_path):
    from spatial_light_modulator_module_b200 import display_holograms as dh, generate_hologram as gh, wavefront_correction as wfc
    g = golden("analytic")
    sub = (slice(None, None, 16), slice(None, None, 16))
    np.testing.assert_array_equal(wfc.deflect_2pi((1.0, 2.0))[sub], g["deflect_sub"])
    ln = gh.lens(0.5, (768, 1024))
    assert ln.dtype == np.uint8
    np.testing.assert_array_equal(ln[sub], g["lens_sub"])
    h0 = np.random.default_rng(int(g["h0_seed"])).uniform(-np.pi, np.pi, size=(768, 1024))
    np.testing.assert_array_equal(gh.deflect_hologram(h0, (1.0, 2.0)), P.deflect_hologram(h0, (1.0, 2.0)))
    np.testing.assert_array_equal(gh.add_lens(h0, 0.5), P.add_lens(h0, 0.5))
    with pytest.raises(ValueError):
        gh.deflect_hologram(np.zeros((128, 128)), (1.0, 2.0))
    q = golden("quantize_96x128")
    p = tmp_path / "h.npy"
    np.save(p, q["hologram"])
    np.testing.assert_array_equal(np.array(dh.mask_hologram(str(p), q["mask"], 256)), q["q2_rand_256"])
    np.testing.assert_array_equal(wfc.convert_2pi_hologram_to_int_hologram(q["hologram_edge"], 255), q["q1_edge_255"])
    np.testing.assert_array_equal(dh.hologram_to_grey(q["hologram"], q["mask"], 200), q["q3_rand_200"])
    from PIL import Image as im
    pp = tmp_path / "g.png"
    im.fromarray(q["png"]).save(pp)
    np.testing.assert_array_equal(np.array(dh.mask_hologram(str(pp), q["mask"], 200)), q["q2png_200"])


def test_sequence_results_pinned_and_pageable_agree(monkeypatch):
    """The movie driver reads a batch back while the next one iterates: into pooled page-locked arrays (plain DMA)
    or, beyond SLM_PINNED_RESULT_BYTES, into ordinary memory from a worker thread.  Same bits either way."""
    from spatial_light_modulator_module_b200 import generate_hologram_sequence as ghs
    frames = synthetic.movie_frames(7, rescale_parameter=5.0)
    h1, e1, err1, _ = ghs.sequence_holograms(frames, 5, precision="fp32", batch=3, want_expected=True)
    monkeypatch.setenv("SLM_PINNED_RESULT_BYTES", "0")
    h2, e2, err2, _ = ghs.sequence_holograms(frames, 5, precision="fp32", batch=3, want_expected=True)
    np.testing.assert_array_equal(h1, h2)
    np.testing.assert_array_equal(e1, e2)
    for a, b in zip(err1, err2):
        np.testing.assert_array_equal(a, b)


def test_sequence_driver_files(tmp_path, monkeypatch):
    """generate_hologram_sequence: PNG frames in, .npy holograms + preview PNGs out
    (generate_hologram_sequence.py:10-32)."""
    from PIL import Image as im
    from spatial_light_modulator_module_b200 import generate_hologram_sequence as ghs
    monkeypatch.chdir(tmp_path)
    src = tmp_path / "images" / "moving_traps" / "seq"
    src.mkdir(parents=True)
    frames = synthetic.movie_frames(5, rescale_parameter=9.0)
    for i, f in enumerate(frames):
        im.fromarray(f).save(src / f"{i}.png")
    a = ns(source_dir="seq", version="v1", max_loops=6, preview=True, precision="fp64", batch=2)
    errors = quiet(ghs.generate_hologram_sequence, a)
    assert len(errors) == 5 and all(len(e) == 6 for e in errors)
    eng = make_engine((768, 1024), "fp64", 1)
    for i in (0, 4):
        h = np.load(tmp_path / "holograms" / "seq_v1_holograms" / f"{i}.npy")
        r = eng.gs(frames[i], 6)
        np.testing.assert_array_equal(h, eng.to_host(r.hologram)[0])
        assert (tmp_path / "images" / "moving_traps" / "seq_v1_preview" / f"{i}.png").exists()
    eng.close()


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_sequence_warm_start(precision):
    """SURVEY 8 f-4, opt-in: frame n starts from frame n-1's hologram.  Equals the hand-made chain of single GS runs
    (the CPU suite also shows the head start it gives on a slowly moving pattern, tests/test_emulated_kernels.py)."""
    from spatial_light_modulator_module_b200 import generate_hologram_sequence as ghs
    frames = synthetic.movie_frames(4, rescale_parameter=0.2)
    holos, _, errors, _ = ghs.sequence_holograms(frames, 6, precision=precision, warm_start=True, gather=False)
    eng = make_engine(frames.shape[1:], precision, 2)     # (a context for more than one plane, like the driver's: same kernels, same bits)
    phasor = None
    for i in range(4):
        r = eng.gs(frames[i], 6, phasor0=phasor)
        np.testing.assert_array_equal(holos[i], eng.to_host(r.hologram)[0])
        np.testing.assert_array_equal(errors[i], r.errors[0])
        phasor = eng.phase_phasor(r.hologram)
    assert all(len(e) == 6 and e[-1] <= e[0] for e in errors[1:])        # the chained frames are valid GS runs
    eng.close()


def test_sequence_uint8_frames_and_writer_callback():
    """8-bit SLM frames (mask add + floor quantisation) straight from the movie driver, device-rasterised
    trap targets, batches handed to a writer callback while later ones iterate -- equal to quantising the float64
    holograms of the host-frame path."""
    from spatial_light_modulator_module_b200 import display_holograms as dh, generate_hologram_sequence as ghs
    n, shape = 9, (768, 1024)
    mask = synthetic.random_mask(shape, seed=3)
    seen = []
    frames8, _, errs8, _ = ghs.sequence_holograms(None, 5, precision="fp32", batch=4, output="uint8", mask=mask, ct2pi=200,
                                                  trap_dots=(synthetic.movie_frame_dots(n, rescale_parameter=7.0), n, shape),
                                                  on_batch=lambda a, b, h, e: seen.append((a, b, h.copy())))
    holos, _, errs, _ = ghs.sequence_holograms(synthetic.movie_frames(n, rescale_parameter=7.0), 5, precision="fp32", batch=4)
    assert frames8.dtype == np.uint8 and [(a, b) for a, b, _ in seen] == [(0, 4), (4, 8), (8, 9)]
    for i in range(n):
        np.testing.assert_array_equal(frames8[i], dh.hologram_to_grey(holos[i], mask, 200))
        np.testing.assert_array_equal(errs8[i], errs[i])
    for a, b, h in seen:
        np.testing.assert_array_equal(h, frames8[a:b])


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_random_phasor_guess(golden, precision):
    pc.check_random_phasor_guess(make_engine, golden, precision)


def test_single_trap_and_device_frames(golden):
    pc.check_single_trap_and_frames(make_engine, golden)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_gif_snapshots(tmp_path, precision):
    from spatial_light_modulator_module_b200 import algorithms
    pc.check_gif_snapshots(None, precision, tmp_path,
                           lambda t, a: quiet(algorithms.gerchberg_saxton, t, a),
                           lambda t, a: quiet(algorithms.gradient_descent, t, a), ns)


def test_move_traps_update_hologram(golden):
    from spatial_light_modulator_module_b200 import move_traps
    g = golden("preview_trap")
    r, c = [int(v) for v in g["trap_rc"]]
    img = np.zeros((192, 256), dtype=np.uint8)
    h = move_traps.update_hologram(img, [(r, c), (5, 5)], 0)
    assert pc.circ(h, g["trap_phase"]).max() < 1e-11
    with pytest.raises(IndexError):
        move_traps.update_hologram(img, [(500, 1)], 0)


# ---- one large plane, slab-decomposed (BASELINE config 5) ------------------------------------------------
@pytest.mark.parametrize("n,precision", [(8192, "fp32"), (16384, "fp32"), (8192, "fp64")])
def test_long_line_transforms(n, precision):
    from scipy.fft import fft, ifft
    from spatial_light_modulator_module_b200.slab import SlabEngine
    rows = 64
    eng = SlabEngine(n, n // rows, 0, precision)
    rng = np.random.default_rng(2)
    x = (rng.standard_normal((rows, n)) + 1j * rng.standard_normal((rows, n))).astype(eng.complex_dtype)
    tol = 3e-6 if precision == "fp32" else 5e-14
    xd, out = eng._mem_upload(x), eng._mem_empty((rows, n), eng.complex_dtype)
    for inverse in (False, True):
        eng._rows_fft(xd, out, inverse)
        ref = (ifft(x.astype(np.complex128), axis=1) * n) if inverse else fft(x.astype(np.complex128), axis=1)
        assert np.abs(eng.to_host(out) - ref).max() / np.abs(ref).max() < tol
    eng.close()


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_slab_gs_single_rank_equals_plane_engine(precision):
    from spatial_light_modulator_module_b200.slab import SlabEngine
    # (fp32 columns of 1024 points run on the warp-per-column kernel, whose butterflies are ordered differently
    #  from the row kernels the slab path uses: bit equality is checked where both use the same line transform)
    n = 1024 if precision == "fp64" else 512
    t = synthetic.shapes_target((n, n))
    ref = make_engine((n, n), precision, 1)
    r = ref.gs(t, 6)
    eng = SlabEngine(n, 1, 0, precision)
    h, e, errs = eng.gs(t, 6)
    # same arithmetic per element; only the order of the error sums differs
    np.testing.assert_array_equal(h, ref.to_host(r.hologram)[0])
    np.testing.assert_allclose(e, ref.to_host(r.expected)[0], rtol=1e-12)
    assert np.max(np.abs(np.array(errs) - r.errors[0]) / r.errors[0]) < (1e-5 if precision == "fp32" else 1e-12)
    eng.close(); ref.close()


@pytest.mark.parametrize("precision,n", [("fp32", 512), ("fp64", 1024), ("fp32", 2048), ("fp64", 2048)])
def test_slab_gd_single_rank_equals_plane_engine(precision, n):
    """GD on the slab path (algorithms.py:60-112 with the plane's max found in a pass of its own) against the ordinary
    engine and the oracle; a tolerance ends both loops on the same iteration."""
    from oracle import numpy_port as P
    from spatial_light_modulator_module_b200 import host_logic as hl
    from spatial_light_modulator_module_b200.slab import SlabEngine
    loops = 6
    t = synthetic.noise_target((n, n), seed=8)
    x0 = hl.host_initial_guess("random", (n, n), 42)
    during, _ = hl.learning_rate_schedule(0.01, 1, loops)
    ref = make_engine((n, n), precision, 1)
    r, _ = ref.gd(t, x0.copy(), during, loops)
    eng = SlabEngine(n, 1, 0, precision)
    h, e, errs = eng.gd(t, x0, during, loops)
    # the ordinary engine's Fourier-plane forms order the scale (norm / max) and the transform's butterflies differently:
    # equal to rounding, not to the bit
    f64 = precision == "fp64"
    ref_h_np, ref_e_np, ref_errs, _ = P.gd_run(t, loops, learning_rate=0.01, unsettle=1)
    for other_h, other_e, other_errs in ((ref.to_host(r.hologram)[0], ref.to_host(r.expected)[0], r.errors[0]),
                                         (ref_h_np, ref_e_np, np.array(ref_errs))):
        assert np.abs(np.angle(np.exp(1j * (h - other_h)))).max() < (1e-9 if f64 else 5e-3)
        np.testing.assert_allclose(e, other_e, rtol=1e-9 if f64 else 1e-3, atol=1e-9 if f64 else 1e-2)
        assert np.max(np.abs(np.array(errs) - other_errs) / other_errs) < (1e-10 if f64 else 1e-4)
    stop_at = float(0.5 * (errs[1] + errs[2]))
    h2, _, errs2 = eng.gd(t, x0, during, loops, tolerance=stop_at)
    assert len(errs2) == 3 and errs2 == errs[:3]
    eng.close(); ref.close()


def test_slab_gs_with_illumination_plane():
    """A non-uniform illumination amplitude on the slab path: the ordinary engine's result (fp64: the same bits)."""
    from spatial_light_modulator_module_b200.slab import SlabEngine
    n = 1024
    t = synthetic.shapes_target((n, n))
    yy, xx = np.mgrid[0:n, 0:n]
    inc = np.exp(-((yy - n / 2) ** 2 + (xx - n / 2) ** 2) / (2 * (n / 3) ** 2))
    ref = make_engine((n, n), "fp64", 1)
    r = ref.gs(t, 5, inc_amp=inc)
    eng = SlabEngine(n, 1, 0, "fp64")
    h, e, errs = eng.gs(t, 5, inc_amp_slab=inc)
    np.testing.assert_array_equal(h, ref.to_host(r.hologram)[0])
    np.testing.assert_allclose(e, ref.to_host(r.expected)[0], rtol=1e-12)
    assert np.max(np.abs(np.array(errs) - r.errors[0]) / r.errors[0]) < 1e-12
    eng.close(); ref.close()


def _golden_slab(name):
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))


def test_slab_gd_8192_vs_reference_golden():
    """GD on a plane only the slab path holds as lines, against the reference's own run (oracle/make_golden.py --slab):
    the error curve (which RISES first on this sparse target, as the reference's does), hologram and expected samples."""
    from spatial_light_modulator_module_b200 import algorithms, slab
    g = _golden_slab("gd_traps_8192_slab")
    n = 8192
    t = synthetic.traps_target((n, n), [(1000, 2000), (6000, 5000), (4096, 700)])
    h, e, errs, _ = slab.gradient_descent_slab(t, 5, learning_rate=0.005, precision="fp32")
    assert np.all(np.abs(h) <= np.pi) and abs(e.max() - 255.0) < 1e-9
    assert abs(algorithms.error_f(e, t, t.size) - errs[-1]) < 1e-5 * errs[-1]
    assert np.max(np.abs(np.array(errs) - g["errors"]) / g["errors"]) < 2e-4
    sub = (slice(None, None, 64), slice(None, None, 64))
    assert np.abs(np.angle(np.exp(1j * (h[sub] - g["hologram_sub"])))).max() < 5e-3
    np.testing.assert_allclose(e[sub], g["expected_sub"], rtol=2e-3, atol=2e-2)


def test_slab_gs_8192_invariants():
    """A plane only the slab path can hold as lines (8192 points): GS on a trap target converges to a fixed
    point, |hologram| <= pi, max(expected) == max(target), error == error_f(expected, target)."""
    from spatial_light_modulator_module_b200 import algorithms
    from spatial_light_modulator_module_b200.slab import SlabEngine
    n = 8192
    t = synthetic.traps_target((n, n), [(1000, 2000), (6000, 5000), (4096, 700)])
    eng = SlabEngine(n, 1, 0, "fp32")
    h, e, errs = eng.gs(t, 6)
    assert np.all(np.abs(h) <= np.pi) and abs(e.max() - 255.0) < 1e-9
    assert abs(algorithms.error_f(e, t, t.size) - errs[-1]) < 1e-5 * errs[-1]
    assert abs(errs[-1] - errs[-2]) < 1e-4 * errs[-1]
    g = _golden_slab("gs_traps_8192_slab")                     # the reference's own run of this case
    assert np.max(np.abs(np.array(errs) - g["errors"]) / g["errors"]) < 2e-4
    sub = (slice(None, None, 64), slice(None, None, 64))
    print("8192 GS hologram vs reference, max circular distance:", np.abs(np.angle(np.exp(1j * (h[sub] - g["hologram_sub"])))).max())
    np.testing.assert_allclose(e[sub], g["expected_sub"], rtol=2e-3, atol=2e-2)
    eng.close()


def test_reused_host_arrays_are_page_locked_safely():
    """Host arrays handed over a second time are page-locked in place (engine._page_lock_if_reused); overlapping
    ranges cannot be locked twice -- that must fall back to the staged copy without leaving a CUDA error behind."""
    from spatial_light_modulator_module_b200 import _ffi as F
    eng = make_engine((1024, 1024), "fp32", 2)
    rng = np.random.default_rng(5)
    base = rng.random((3, 1024, 1024)) * 2 * np.pi
    mask = rng.random((1024, 1024)) * 2 * np.pi
    ref = {}
    for sl in (slice(0, 2), slice(1, 3)):
        ref[sl.start] = ((base[sl] + mask) % (2 * np.pi) * 256 / (2 * np.pi)).astype(np.uint8)
    for rep in range(3):                     # 2nd use of [0:2] locks it; [1:3] overlaps it and must still work
        for sl in (slice(0, 2), slice(1, 3)):
            out = eng.to_host(eng.quantize(base[sl], mask, 256, F.QUANT_FLOOR))
            np.testing.assert_array_equal(out, ref[sl.start])
    base[0, 0, 0] = 1.0                      # the locked array is still ordinary, writable memory
    assert eng.to_host(eng.quantize(base[0:2], mask, 256, F.QUANT_FLOOR))[0, 0, 0] == np.uint8(((1.0 + mask[0, 0]) % (2 * np.pi)) * 256 / (2 * np.pi))
    eng.close()


def test_device_mt19937_stream():
    pc.check_device_mt19937(make_engine)


def test_batched_algorithm_comparison_matches_single_calls():
    """Config-4 shape at reduced size: batched GD + GS curves equal the one-target-at-a-time drop-in calls."""
    from spatial_light_modulator_module_b200 import algorithms, compare_error_evolution_algorithms as cmp
    shape = (256, 256)
    targets = np.stack([synthetic.noise_target(shape, seed=1), synthetic.shapes_target(shape), synthetic.traps_target(shape)])
    a = ns(max_loops=8, precision="fp64")
    cmp.fill_unnecessary_args(a)
    gd, gs = cmp.error_evolution_curves(targets, a, batch=2)
    assert len(gd) == len(gs) == 3
    for i, t in enumerate(targets):
        b = ns(max_loops=8, precision="fp64")
        _, _, e_gd = quiet(algorithms.gradient_descent, t, b)
        _, _, e_gs = quiet(algorithms.gerchberg_saxton, t, b)
        np.testing.assert_allclose(gd[i], np.array(e_gd), rtol=1e-12)
        np.testing.assert_allclose(gs[i], np.array(e_gs), rtol=1e-12)
