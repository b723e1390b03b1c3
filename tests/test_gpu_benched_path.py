"""Parity of the code path `bench.py` times: large batches at 1024x1024 / 768x1024, where the
engine leaves the few-planes forms (in-kernel closing, one fused Fourier-plane pass per GD iteration) for the
large-batch ones; the column kernels of 2048- and 4096-point lines; the 8-bit frame tolerance of the north star.

Every test goes through the C ABI (Engine -> ctypes -> lib/libslmholo.so) and compares with the CPU oracle
(oracle/numpy_port.py, pinned bit-exact to the reference) or with fixtures the unmodified reference generated
(oracle/make_golden.py --benched-path).

Tolerances (north star): fp64 error curves 1e-9 relative; fp32 intensity 1e-3 of max, circular phase 1e-3 rad
(on >= 99.9 % of the pixels: phases of near-zero field values are ill conditioned), 8-bit frames +-1 LSB
(mod ct2pi) on at most 2e-3 of the pixels in fp32 (measured 8.7e-4: a phase error of 1e-4 rad is 4e-3 grey levels, so about
that fraction of the pixels sits close enough to a grey-level boundary to cross it) and 1e-6 in fp64.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from oracle import numpy_port as P
from spatial_light_modulator_module_b200 import _ffi, host_logic as hl, synthetic
from tests import parity_checks as pc

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_engine(shape, precision, max_batch):
    from spatial_light_modulator_module_b200.engine import Engine
    return Engine(shape, precision, max_batch)


def plane_targets(shape, n):
    """n different targets: noise (different seeds), the shapes target, a trap target."""
    out = [synthetic.noise_target(shape, seed=100 + i) for i in range(n)]
    if n > 2:
        out[2] = synthetic.shapes_target(shape)
    if n > 5:
        out[5] = synthetic.traps_target(shape, [(shape[0] // 3, shape[1] // 4), (2 * shape[0] // 3, shape[1] // 2)])
    return np.stack(out)


def guesses(shape, n):
    return np.stack([hl.host_initial_guess("random", shape, 40 + i) for i in range(n)])


# ---- GD, batch >= 8: max pass that keeps the transform + gradient pass (or the pipelined one-pass form), planes closed
# ---- by the kernel behind the pass -----------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("shape", [(1024, 1024), (768, 1024)])
def test_gd_large_batch_vs_oracle_and_single_runs(shape, precision):
    batch, loops = 8, 10
    tol = pc.TOL[precision]
    t, x0 = plane_targets(shape, batch), guesses(shape, batch)
    during, _ = hl.learning_rate_schedule(0.005, 0, loops)
    eng = make_engine(shape, precision, batch)
    res, _ = eng.gd(t, x0.copy(), during, loops)
    holo, exp = eng.to_host(res.hologram), eng.to_host(res.expected)
    # every plane equals its single-plane run through the few-planes code path: same arithmetic, same bits
    for b in (1, 2, 5, 7):
        r1, _ = eng.gd(t[b], x0[b].copy(), during, loops)
        np.testing.assert_array_equal(r1.errors[0], res.errors[b])
        np.testing.assert_array_equal(eng.to_host(r1.hologram)[0], holo[b])
        np.testing.assert_array_equal(eng.to_host(r1.expected)[0], exp[b])
    # ... and the oracle (algorithms.py:83-105) on three of the planes
    for b in (0, 2, 5):
        ref_h, ref_e, ref_errs, _ = P.gd_run(t[b], loops, seed=40 + b)
        assert len(res.errors[b]) == loops
        assert np.max(np.abs(res.errors[b] - np.array(ref_errs)) / np.abs(ref_errs)) < tol["gd_curve"]
        assert np.mean(pc.circ(holo[b], ref_h) < tol["gd_phase"]) >= 0.999
        assert np.abs(exp[b] - ref_e).max() <= tol["gd_inten"] * ref_e.max()
    eng.close()


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_gd_large_batch_tolerance_stops_planes_independently(precision):
    """tolerance > 0 in a large batch: planes leave the loop one by one (algorithms.py:83) and still equal their
    single-plane runs bit for bit."""
    shape, batch, loops = (1024, 1024), 8, 8
    t, x0 = plane_targets(shape, batch), guesses(shape, batch)
    during, _ = hl.learning_rate_schedule(0.005, 0, loops)
    eng = make_engine(shape, precision, batch)
    free, _ = eng.gd(t, x0.copy(), during, loops)
    e2 = free.errors[2]                                           # the shapes plane: errors fall monotonically
    tol = float(0.5 * (e2[3] + e2[4]))
    res, _ = eng.gd(t, x0.copy(), during, loops, tol)
    holo = eng.to_host(res.hologram)
    stops = []
    for b in range(batch):
        ref = free.errors[b]
        below = ~(ref > tol)
        stop = int(np.argmax(below)) + 1 if np.any(below) else loops
        stops.append(stop)
        assert len(res.errors[b]) == stop == res.iterations[b]
        np.testing.assert_array_equal(res.errors[b], ref[:stop])
    assert min(stops) < loops and max(stops) == loops
    for b in (2, int(np.argmax(stops))):
        r1, _ = eng.gd(t[b], x0[b].copy(), during[:stops[b]], stops[b])
        np.testing.assert_array_equal(eng.to_host(r1.hologram)[0], holo[b])
    eng.close()


# ---- GS, batch >= 8 --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("shape", [(1024, 1024), (768, 1024)])
def test_gs_large_batch_teacher_forced_and_single_runs(shape, precision, monkeypatch):
    """Ten GS iterations, each started on the device from the oracle's state entering it (GS free-running is chaotic
    on dense targets), all eight planes in one batch; then a free-running batch against single-plane
    runs in the same context, bit for bit; a context made for one plane (plain column kernel) within tolerances."""
    batch, steps = 8, 10
    tol = pc.TOL[precision]
    t = plane_targets(shape, batch)
    eng = make_engine(shape, precision, batch)
    checked = (0, 2, 5)
    st = {b: P.gs_setup(t[b]) for b in checked}
    B0 = {b: P.gs_first_phasor(st[b]) for b in checked}
    filler = np.exp(2j * np.pi * np.random.default_rng(3).random(shape))
    for k in range(steps):
        phasors = np.stack([filler] * batch)
        refs = {}
        for b in checked:
            phasors[b] = B0[b] if k == 0 else st[b].inc_amp * P.unit_phasor(st[b].A)
            _, exp_ref, err_ref = P.gs_step(st[b])
            refs[b] = (exp_ref, err_ref, st[b].A.copy())
        if k not in (0, 1, 4, 9):
            continue
        res = eng.gs(t, 1, phasor0=phasors)
        exp, holo = eng.to_host(res.expected), eng.to_host(res.hologram)
        for b in checked:
            exp_ref, err_ref, A = refs[b]
            assert abs(res.errors[b][0] - err_ref) <= tol["err"] * abs(err_ref)
            assert np.abs(exp[b] - exp_ref).max() <= tol["inten"] * exp_ref.max()
            w = np.abs(A) / np.abs(A).max()
            assert (pc.circ(holo[b], np.angle(A)) * w).max() < tol["wphase"]
    res = eng.gs(t, 10)
    holo, exp = eng.to_host(res.hologram), eng.to_host(res.expected)
    for b in (0, 3, 5, 7):
        r1 = eng.gs(t[b], 10)
        np.testing.assert_array_equal(r1.errors[0], res.errors[b])
        np.testing.assert_array_equal(eng.to_host(r1.hologram)[0], holo[b])
        np.testing.assert_array_equal(eng.to_host(r1.expected)[0], exp[b])
    # a context for ONE plane takes the plain column kernel (shorter latency, another rounding): the same iteration;
    # SLM_GS_GROUPS=1 keeps it on the pipeline kernels and gives the batch's bits
    one = make_engine(shape, precision, 1)
    r1 = one.gs(t[3], 10)
    assert abs(r1.errors[0][0] - res.errors[3][0]) <= tol["err"] * res.errors[3][0]
    assert abs(r1.errors[0][1] - res.errors[3][1]) <= (1e-2 if precision == "fp32" else 1e-8) * res.errors[3][1]   # (chaotic growth)
    monkeypatch.setenv("SLM_GS_GROUPS", "1")
    r1 = one.gs(t[3], 10)
    np.testing.assert_array_equal(r1.errors[0], res.errors[3])
    np.testing.assert_array_equal(one.to_host(r1.hologram)[0], holo[3])
    one.close()
    eng.close()


def test_large_batch_closing_forms_agree():
    """The planes of a large batch are closed by a kernel behind the pass; SLM_NO_DEFER_CLOSE=1 keeps the in-kernel
    closing, SLM_NO_FUSED_GD=1 the two-pass GD form.  Same bits for GS and GD (child processes: the switches are read
    once per process)."""
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from spatial_light_modulator_module_b200.engine import Engine\n"
        "from spatial_light_modulator_module_b200 import synthetic, host_logic as hl\n"
        "shape=(1024,1024); n=8; eng=Engine(shape,'fp32',n)\n"
        "t=np.stack([synthetic.noise_target(shape,seed=i) for i in range(n)])\n"
        "x0=np.stack([hl.host_initial_guess('random',shape,42+i) for i in range(n)])\n"
        "during,_=hl.learning_rate_schedule(0.005,0,6)\n"
        "r,_=eng.gd(t,x0,during,6)\n"
        "g=eng.gs(t,6)\n"
        "np.savez(sys.argv[1], e=np.array(r.errors), h=eng.to_host(r.hologram), x=eng.to_host(r.expected),\n"
        "         ge=np.array(g.errors), gh=eng.to_host(g.hologram))\n"
    ) % ROOT
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, env in (("default", {}), ("in_kernel_close", {"SLM_NO_DEFER_CLOSE": "1"}), ("two_pass", {"SLM_NO_FUSED_GD": "1"})):
            path = os.path.join(tmp, name + ".npz")
            subprocess.run([sys.executable, "-c", code, path], check=True, env={**os.environ, **env}, timeout=900)
            out[name] = dict(np.load(path))
    for other in ("in_kernel_close", "two_pass"):
        for k in ("e", "h", "x", "ge", "gh"):
            np.testing.assert_array_equal(out["default"][k], out[other][k], err_msg=f"{other}:{k}")


# ---- the benched configuration itself against the unmodified reference ---------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_gd_bench_configuration_vs_reference_golden(golden, precision):
    """bench.py's default step: batch 32 x 1024x1024, 100 iterations, fp32 (and the same in fp64, batch 8).  Plane 0
    carries the target and the seed-42 random guess of the reference-generated fixture
    gd_noise_1024x1024_curves.npz (100 iterations of /root/reference/src/algorithms.py:60-112)."""
    g = golden("gd_noise_1024x1024_curves")
    shape, loops = (1024, 1024), 100
    batch = 32 if precision == "fp32" else 8
    t = np.stack([synthetic.noise_target(shape, seed=(0 if i == 0 else 500 + i)) for i in range(batch)])
    x0 = np.stack([hl.host_initial_guess("random", shape, 42)] * batch)
    during, _ = hl.learning_rate_schedule(0.005, 0, loops)
    eng = make_engine(shape, precision, batch)
    res, _ = eng.gd(t, x0, during, loops)
    e = res.errors[0]
    assert len(e) == loops
    ctol, ptol, itol = (1e-9, 1e-8, 1e-9) if precision == "fp64" else (1e-3, 1e-3, 1e-3)
    assert np.max(np.abs(e - g["errors"]) / g["errors"]) < ctol
    holo = eng.to_host(res.hologram[0:1])[0]
    exp = eng.to_host(res.expected[0:1])[0]
    sub = (slice(None, None, 16), slice(None, None, 16))
    assert np.mean(pc.circ(holo[sub], g["hologram_sub"]) < ptol) >= 0.999
    assert np.abs(exp[sub] - g["expected_sub"]).max() < itol * g["expected_sub"].max()
    # the other planes ran the same loop: every curve falls, none equals plane 0's
    for b in range(1, batch):
        assert len(res.errors[b]) == loops and res.errors[b][-1] < res.errors[b][0]
        assert res.errors[b][0] != e[0]
    # 8-bit frame of plane 0 (mask add + floor quantise, move_traps.py:135-140) against the reference's frame
    mask = synthetic.random_mask(shape, seed=1)
    frame = eng.to_host(eng.quantize(res.hologram[0:1], mask, 256, _ffi.QUANT_FLOOR))[0]
    check_frame(frame[::4, ::4], g["q3_sub"], 256, 2e-3 if precision == "fp32" else 1e-6)
    eng.close()


def check_frame(frame, ref, ct2pi, max_fraction):
    """+-1 LSB modulo ct2pi on at most `max_fraction` of the pixels, identical elsewhere."""
    d = (frame.astype(np.int32) - ref.astype(np.int32)) % ct2pi
    d = np.minimum(d, ct2pi - d)
    assert d.max() <= 1, f"frames differ by {d.max()} grey levels"
    assert np.mean(d != 0) <= max_fraction, f"{np.mean(d != 0):.2e} of the pixels differ"


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_config2_quantised_frames_within_one_lsb(golden, precision):
    """North star: 'quantised 8-bit holograms agree with at most +-1 LSB on a stated pixel fraction'.  Config 2 at the
    SLM shape through the drop-in API: gradient_descent (100 iterations) -> mask add -> Q3 (floor) and Q2 (PIL float32
    path) frames against the frames of the reference's own hologram.  Stated fraction: 2e-3 (fp32), 1e-6 (fp64)."""
    import argparse
    import contextlib
    import io
    from spatial_light_modulator_module_b200 import algorithms, display_holograms as dh
    g = golden("gd_noise_768x1024_frames")
    shape = (768, 1024)
    t = synthetic.noise_target(shape, seed=0)
    mask = synthetic.random_mask(shape, seed=1)
    a = argparse.Namespace(incomming_intensity="uniform", tolerance=0, max_loops=100, gif=False, print_info=False,
                           plot_error=False, initial_guess="random", random_seed=42, white_attention=1,
                           learning_rate=0.005, unsettle=0, precision=precision)
    with contextlib.redirect_stdout(io.StringIO()):
        holo, _, errs = algorithms.gradient_descent(t, a)
    frac = 2e-3 if precision == "fp32" else 1e-6
    check_frame(dh.hologram_to_grey(holo, mask, 256)[::4, ::4], g["q3_sub"], 256, frac)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "h.npy")
        np.save(path, holo)
        q2 = np.array(dh.mask_hologram(path, mask, 256))
    check_frame(q2[::4, ::4], g["q2_sub"], 256, frac)


def test_gs_1024_first_iterations_vs_reference_golden(golden):
    """GS at the metric shape from the device's own setup (complex64 first ifft2): the first iterations
    match the reference's curve; later ones diverge chaotically and are compared statistically."""
    g = golden("gs_noise_1024x1024_curves")
    t = synthetic.noise_target((1024, 1024), seed=0)
    for precision in ("fp32", "fp64"):
        eng = make_engine((1024, 1024), precision, 1)
        res = eng.gs(t, 10)
        e = res.errors[0]
        assert abs(e[0] - g["errors"][0]) < 1e-5 * g["errors"][0]
        assert abs(e[1] - g["errors"][1]) < 1e-2 * g["errors"][1]      # (measured 1.3e-3: the setup field is complex64 on both sides)
        # iterations 2..6 sit on the plateau of the Hermitian-symmetric manifold (C = fft2(B) stays real, SURVEY finding 2);
        # WHEN rounding noise breaks the symmetry differs between FFT implementations (the reference leaves the plateau at
        # iteration 7 here, upwards), so later errors are only required to stay in the range a GS run visits
        assert np.all(e > 0.2 * g["errors"].min()) and np.all(e < 1.5 * g["errors"].max())
        eng.close()


# ---- 2048- and 4096-point columns: the column-group kernel and the one-CTA-per-tile kernel ---------------------------
@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("shape", [(2048, 2048), (4096, 4096), (4096, 2048)])
def test_long_columns_teacher_forced(shape, precision):
    """One GS iteration from the oracle's state entering iterations 0 and 1, and two free GD iterations, at the plane
    sizes of configs 4 and the north star's upper end."""
    tol = pc.TOL[precision]
    t = synthetic.noise_target(shape, seed=11)
    eng = make_engine(shape, precision, 2)
    st = P.gs_setup(t)
    for k in range(2):
        B = P.gs_first_phasor(st) if k == 0 else st.inc_amp * P.unit_phasor(st.A)
        _, exp_ref, err_ref = P.gs_step(st)
        res = eng.gs(np.stack([t, t[::-1]]), 1, phasor0=np.stack([B, B[::-1]]))
        assert abs(res.errors[0][0] - err_ref) <= tol["err"] * abs(err_ref)
        exp = eng.to_host(res.expected[0:1])[0]
        assert np.abs(exp - exp_ref).max() <= tol["inten"] * exp_ref.max()
        holo = eng.to_host(res.hologram[0:1])[0]
        w = np.abs(st.A) / np.abs(st.A).max()
        assert (pc.circ(holo, np.angle(st.A)) * w).max() < tol["wphase"]
        del exp, holo, res
    loops = 2
    x0 = hl.host_initial_guess("random", shape, 42)
    during, _ = hl.learning_rate_schedule(0.005, 0, loops)
    res, _ = eng.gd(np.stack([t, t[::-1]]), np.stack([x0, x0[::-1]]), during, loops)
    ref_h, ref_e, ref_errs, _ = P.gd_run(t, loops)
    assert np.max(np.abs(res.errors[0] - np.array(ref_errs)) / np.abs(ref_errs)) < tol["gd_curve"]
    assert np.mean(pc.circ(eng.to_host(res.hologram[0:1])[0], ref_h) < tol["gd_phase"]) >= 0.999
    assert np.abs(eng.to_host(res.expected[0:1])[0] - ref_e).max() <= tol["gd_inten"] * ref_e.max()
    eng.close()
