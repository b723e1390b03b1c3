"""Parity checks shared by the CPU suite (host emulation of the kernel sources) and the GPU suite
(the real library on a B200).  Each check takes ``make_engine(shape, precision, max_batch)`` and
compares the engine with the oracle (oracle/numpy_port.py) or the committed golden fixtures.

Tolerances (stated once, used by both suites):
    fp64  teacher-forced step: error rel 1e-12, intensity 1e-12 of max, amplitude-weighted phase 1e-10 rad
          free-running GD: error curve rel 1e-9, phase 1e-8 rad
    fp32  teacher-forced step: error rel 1e-5, intensity 1e-5 of max (north star: 1e-3),
          amplitude-weighted phase 1e-4 rad (north star: 1e-3)
          free-running GD: error curve rel 1e-4, intensity 1e-4 of max, phase 1e-3 rad on >= 99.9 % of pixels
    analytic holograms and quantisers: bit exact
"""
from __future__ import annotations

import numpy as np
import pytest
from scipy.fft import fft2, ifft2

from oracle import numpy_port as P
from spatial_light_modulator_module_b200 import _ffi, constants as c, host_logic as hl, synthetic

TOL = {
    "fp64": dict(fft=2e-14, err=1e-12, inten=1e-12, wphase=1e-10, gd_curve=1e-9, gd_phase=1e-8, gd_inten=1e-10),
    "fp32": dict(fft=2e-6, err=1e-5, inten=1e-5, wphase=1e-4, gd_curve=1e-4, gd_phase=1e-3, gd_inten=1e-4),
}


def circ(a, b):
    return np.abs(np.angle(np.exp(1j * (a - b))))


def targets(shape):
    return {"noise": synthetic.noise_target(shape, seed=3), "shapes": synthetic.shapes_target(shape),
            "traps": synthetic.traps_target(shape)}


# ---------------------------------------------------------------------------------------------
def check_fft2(make_engine, shape, precision, batch=2):
    eng = make_engine(shape, precision, batch)
    rng = np.random.default_rng(7)
    x = (rng.standard_normal((batch,) + shape) + 1j * rng.standard_normal((batch,) + shape))
    for inverse in (False, True):
        out = eng.to_host(eng.fft2(x.astype(eng.complex_dtype), inverse))
        ref = (ifft2 if inverse else fft2)(x.astype(eng.complex_dtype).astype(np.complex128), axes=(1, 2))
        assert np.abs(out - ref).max() / np.abs(ref).max() < TOL[precision]["fft"]
    # in place
    xd = eng._mem_upload(x.astype(eng.complex_dtype))
    eng._check(eng._lib.slm_fft2(eng._ctx, batch, eng._mem_ptr(xd), eng._mem_ptr(xd), 0))
    ref = fft2(x.astype(eng.complex_dtype).astype(np.complex128), axes=(1, 2))
    assert np.abs(eng.to_host(xd) - ref).max() / np.abs(ref).max() < TOL[precision]["fft"]
    eng.close()


def check_gs_teacher_forced(make_engine, shape, precision, kind, steps=(0, 1, 4)):
    """State of the oracle entering iteration k -> ONE engine iteration -> compare everything the
    iteration produces (error, expected_outcome, angle of the next field)."""
    tol = TOL[precision]
    t = targets(shape)[kind]
    eng = make_engine(shape, precision, 1)
    st = P.gs_setup(t)
    for k in range(max(steps) + 1):
        B = P.gs_first_phasor(st) if k == 0 else st.inc_amp * P.unit_phasor(st.A)
        _, exp_ref, err_ref = P.gs_step(st)          # advances st.A
        if k not in steps:
            continue
        res = eng.gs(t, 1, phasor0=B)
        assert len(res.errors[0]) == 1
        assert abs(res.errors[0][0] - err_ref) <= tol["err"] * abs(err_ref) + 1e-300
        exp = eng.to_host(res.expected)[0]
        assert np.abs(exp - exp_ref).max() <= tol["inten"] * exp_ref.max()
        holo = eng.to_host(res.hologram)[0]
        w = np.abs(st.A) / np.abs(st.A).max()
        assert (circ(holo, np.angle(st.A)) * w).max() < tol["wphase"]
    eng.close()


def check_gs_device_setup(make_engine, shape, precision, kind, loops=6):
    """Free-running GS from the device's own setup (A = ifft2(amplitude) in complex64).  Only the
    first error is deterministic enough to compare tightly (the later trajectory is chaotic on
    dense targets and noise-seeded at the nodes of trap targets); the rest must stay a
    valid GS run: finite, error curve of the right length, unit-range expected outcome."""
    t = targets(shape)[kind]
    eng = make_engine(shape, precision, 1)
    res = eng.gs(t, loops)
    _, _, errs = P.gs_run(t, loops)
    e = res.errors[0]
    assert len(e) == loops and np.all(np.isfinite(e))
    first_tol = 1e-2 if kind == "traps" else 1e-5
    assert abs(e[0] - errs[0]) <= first_tol * errs[0]
    exp = eng.to_host(res.expected)[0]
    assert abs(exp.max() - float(t.max())) < 1e-6 * float(t.max())
    holo = eng.to_host(res.hologram)[0]
    assert np.all(np.abs(holo) <= np.pi + 1e-12)
    if kind != "traps":
        assert e[-1] < e[0] and 0.5 * errs[-1] < e[-1] < 1.5 * errs[-1]     # same basin statistically
    eng.close()


def run_gd(eng, t, loops, tolerance=0.0, initial_guess="random", seed=42, white_attention=1, learning_rate=0.005,
           unsettle=0):
    x0 = hl.host_initial_guess(initial_guess, t.shape, seed)
    if x0 is None:
        x0 = eng.fourier_guess(t)
    during, after = hl.learning_rate_schedule(learning_rate, unsettle, loops)
    res, _ = eng.gd(t, x0, during, loops, tolerance, white_attention=white_attention)
    return res, after[len(res.errors[0])]


def check_gd_golden(make_engine, golden, name, precision):
    """Free-running GD against the reference's own output (tests/golden, made by the unmodified
    reference)."""
    tol = TOL[precision]
    g = golden(name)
    kw = {}
    for k in g.files:
        if k.startswith("arg_"):
            v = g[k][()]
            kw[k[4:]] = v.item() if hasattr(v, "item") else v
    loops = kw.pop("max_loops")
    tolerance = kw.pop("tolerance", 0.0)
    if "initial_guess" in kw:
        kw["initial_guess"] = str(kw["initial_guess"])
    t = g["target"]
    eng = make_engine(t.shape, precision, 1)
    with np.errstate(all="ignore"):
        res, lr = run_gd(eng, t, loops, tolerance, **kw)
    e, ref = res.errors[0], g["errors"]
    assert len(e) == len(ref)
    assert lr == float(g["final_learning_rate"])
    loose = kw.get("initial_guess") == "fourier"        # starts from the device's complex64 setup field
    ctol = 1e-3 if loose else tol["gd_curve"]
    assert np.max(np.abs(e - ref) / np.abs(ref)) < ctol
    if not loose:
        holo = eng.to_host(res.hologram)[0]
        d = circ(holo, g["hologram"])
        assert np.mean(d < tol["gd_phase"]) >= 0.999
        exp = eng.to_host(res.expected)[0]
        assert np.abs(exp - g["expected"]).max() <= tol["gd_inten"] * g["expected"].max()
    eng.close()


def check_gd_vs_oracle(make_engine, shape, precision, kind="noise", loops=4, batch=1):
    """Free-running GD against the oracle's restatement at an arbitrary plane shape (the golden fixtures only
    hold a few shapes); with ``batch`` > 1 the same target is run in every plane of a batch."""
    tol = TOL[precision]
    t = targets(shape)[kind]
    eng = make_engine(shape, precision, batch)
    x0 = hl.host_initial_guess("random", t.shape, 42)
    during, _ = hl.learning_rate_schedule(0.005, 0, loops)
    tb = np.stack([t] * batch)
    res, _ = eng.gd(tb, np.stack([x0] * batch), during, loops)
    ref_h, ref_e, ref_errs, _ = P.gd_run(t, loops)
    for b in range(batch):
        e = res.errors[b]
        assert len(e) == loops
        assert np.max(np.abs(e - np.array(ref_errs)) / np.abs(ref_errs)) < tol["gd_curve"]
        assert np.mean(circ(eng.to_host(res.hologram)[b], ref_h) < tol["gd_phase"]) >= 0.999
        assert np.abs(eng.to_host(res.expected)[b] - ref_e).max() <= tol["gd_inten"] * ref_e.max()
    eng.close()


def check_gd_tolerance_and_batch(make_engine, precision, shape=(128, 128), loops=8):
    """GD loop condition per plane (algorithms.py:83): the planes of one batch stop independently and every plane
    equals its single-plane run (error curve, hologram, expected outcome) bit for bit."""
    tt = targets(shape)
    order = ["noise", "shapes", "traps"]
    eng = make_engine(shape, precision, 3)
    x0 = hl.host_initial_guess("random", shape, 42)
    during, _ = hl.learning_rate_schedule(0.005, 0, loops)
    singles = {}
    for k in order:
        r, _ = eng.gd(tt[k], x0.copy(), during, loops)
        singles[k] = (r.errors[0], eng.to_host(r.hologram)[0], eng.to_host(r.expected)[0])
    e_mid = singles["shapes"][0]
    tol = float(0.5 * (e_mid[loops // 2 - 1] + e_mid[loops // 2]))      # GD errors fall monotonically here: "shapes" stops midway
    res, _ = eng.gd(np.stack([tt[k] for k in order]), np.stack([x0] * 3), during, loops, tol)
    holo, exp = eng.to_host(res.hologram), eng.to_host(res.expected)
    stops = []
    for i, k in enumerate(order):
        ref_err = singles[k][0]
        below = ~(ref_err > tol)
        stop = int(np.argmax(below)) + 1 if np.any(below) else loops
        stops.append(stop)
        assert len(res.errors[i]) == stop == res.iterations[i]
        np.testing.assert_array_equal(res.errors[i], ref_err[:stop])
        if stop == loops:
            np.testing.assert_array_equal(holo[i], singles[k][1])
            np.testing.assert_array_equal(exp[i], singles[k][2])
        else:                                            # equals a run of exactly `stop` iterations
            r2, _ = eng.gd(tt[k], x0.copy(), during[:stop], stop)
            np.testing.assert_array_equal(eng.to_host(r2.hologram)[0], holo[i])
            np.testing.assert_array_equal(eng.to_host(r2.expected)[0], exp[i])
    assert min(stops) < loops
    eng.close()


def check_gs_tolerance_and_batch(make_engine, precision, shape=(128, 128)):
    """Loop condition per plane (algorithms.py:29): planes of one batch stop independently, and a
    batched run equals the single-plane runs."""
    tt = targets(shape)
    st = {k: P.gs_setup(v) for k, v in tt.items()}
    B0 = {k: P.gs_first_phasor(s) for k, s in st.items()}
    eng = make_engine(shape, precision, 3)
    # single-plane references from the engine itself
    singles = {k: eng.gs(tt[k], 10, 0.0, phasor0=B0[k]) for k in tt}
    single_host = {k: (eng.to_host(r.hologram)[0], eng.to_host(r.expected)[0], r.errors[0]) for k, r in singles.items()}
    order = ["noise", "shapes", "traps"]
    # tolerance: between the first and the last error of the "shapes" run -> it stops early
    e_sh = single_host["shapes"][2]
    tol = float(np.sort(e_sh)[len(e_sh) // 2])
    stop_sh = int(np.argmax(~(e_sh > tol))) + 1
    res = eng.gs(np.stack([tt[k] for k in order]), 10, tol, phasor0=np.stack([B0[k] for k in order]))
    holo, exp = eng.to_host(res.hologram), eng.to_host(res.expected)
    for i, k in enumerate(order):
        ref_err = single_host[k][2]
        stop = int(np.argmax(~(ref_err > tol))) + 1 if np.any(~(ref_err > tol)) else 10
        assert len(res.errors[i]) == stop == res.iterations[i]
        np.testing.assert_array_equal(res.errors[i], ref_err[:stop])
        if stop == 10:
            np.testing.assert_array_equal(holo[i], single_host[k][0])
            np.testing.assert_array_equal(exp[i], single_host[k][1])
    assert stop_sh < 10
    # the early-stopped plane equals a run of exactly that many loops
    r2 = eng.gs(tt["shapes"], stop_sh, 0.0, phasor0=B0["shapes"])
    i = order.index("shapes")
    np.testing.assert_array_equal(eng.to_host(r2.hologram)[0], holo[i])
    np.testing.assert_array_equal(eng.to_host(r2.expected)[0], exp[i])
    eng.close()


def check_gd_tolerance(make_engine, golden, precision):
    check_gd_golden(make_engine, golden, "gd_traps_tol_128x128", precision)


def check_gs_real_targets(make_engine, golden, precision):
    """Non-8-bit targets follow numpy's sqrt dtype rule: first error vs the reference."""
    for name in ("gs_noise_float64_128x128", "gs_noise_float32_128x128", "gs_noise_uint16_128x128"):
        g = golden(name)
        t = g["target"]
        eng = make_engine(t.shape, precision, 1)
        res = eng.gs(t, 3)
        assert abs(res.errors[0][0] - g["errors"][0]) <= 1e-5 * g["errors"][0]
        st = P.gs_setup(t)
        B0 = P.gs_first_phasor(st)
        res = eng.gs(t, 2, phasor0=B0)
        tol = TOL[precision]["err"] * 100
        assert np.max(np.abs(res.errors[0] - g["errors"][:2]) / g["errors"][:2]) < tol
        eng.close()


def check_illumination(make_engine, precision):
    """Non-uniform incomming amplitude multiplies B (algorithms.py:30) and the GD fields (:84,:88).
    With an 8-bit illumination image the reference's amplitude is float16, which keeps its whole
    GS loop in complex64 (float16 * complex64 -> complex64), so the oracle itself is only
    single-precision accurate here: tolerances are 1e-5 in both engine precisions."""
    shape = (128, 128)
    t = synthetic.noise_target(shape, seed=9)
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
    inten = (255 * np.exp(-((yy - 64) ** 2 + (xx - 64) ** 2) / (2 * 40.0 ** 2))).astype(np.uint8)
    inc = np.sqrt(inten).astype(np.float64)      # float16 values, as the reference gets from an 8-bit image
    eng = make_engine(shape, precision, 1)
    st = P.gs_setup(t, inten)
    B0 = P.gs_first_phasor(st)
    _, exp_ref, err_ref = P.gs_step(st)
    res = eng.gs(t, 1, inc_amp=inc, phasor0=B0)
    assert abs(res.errors[0][0] - err_ref) < 1e-5 * err_ref
    # GS second iteration uses inc on the device side
    B1 = st.inc_amp * P.unit_phasor(st.A)
    _, _, err1 = P.gs_step(st)
    res = eng.gs(t, 2, inc_amp=inc, phasor0=B0)
    assert abs(res.errors[0][1] - err1) < 1e-3 * err1
    # GD
    with np.errstate(all="ignore"):
        holo, exp, errs, _ = P.gd_run(t, 6, incomming_intensity=inten)
    x0 = hl.host_initial_guess("random", shape, 42)
    during, _ = hl.learning_rate_schedule(0.005, 0, 6)
    r, _ = eng.gd(t, x0, during, 6, inc_amp=inc)
    assert np.max(np.abs(r.errors[0] - errs) / np.abs(errs)) < max(TOL[precision]["gd_curve"] * 10, 1e-5)
    eng.close()


def check_analytic_and_quantisers(make_engine, golden):
    eng = make_engine((64, 64), "fp64", 1)
    g = golden("analytic")
    sub = (slice(None, None, 16), slice(None, None, 16))
    d = eng.to_host(eng.deflect_phase((1.0, 2.0), c.px_distance, c.wavelength, c.u, (c.slm_height, c.slm_width)))
    np.testing.assert_array_equal(d, P.deflect_phase((1.0, 2.0)))
    np.testing.assert_array_equal(d[sub], g["deflect_sub"])
    d2 = eng.to_host(eng.deflect_phase((-0.5, 0.25), c.px_distance, c.wavelength, c.u, (c.slm_height, c.slm_width)))
    np.testing.assert_array_equal(d2[sub], g["deflect2_sub"])
    ln = eng.to_host(eng.lens_phase(0.5, c.px_distance, c.wavelength, (768, 1024), True))
    np.testing.assert_array_equal(ln.astype(np.uint8)[sub], g["lens_sub"])
    np.testing.assert_array_equal(ln, P.lens_phase(0.5, (768, 1024)).astype(np.float64))
    ln2 = eng.to_host(eng.lens_phase(-1.25, c.px_distance, c.wavelength, (96, 128), True))
    np.testing.assert_array_equal(ln2.astype(np.uint8), g["lens2"])
    lnf = eng.to_host(eng.lens_phase(0.5, c.px_distance, c.wavelength, (96, 128), False))
    np.testing.assert_array_equal(lnf, P.lens_phase(0.5, (96, 128), uint8_quirk=False))
    h0 = np.random.default_rng(int(g["h0_seed"])).uniform(-np.pi, np.pi, size=(768, 1024))
    np.testing.assert_array_equal(eng.to_host(eng.add_mod2pi(h0, d)), P.deflect_hologram(h0, (1.0, 2.0)))
    np.testing.assert_array_equal(eng.to_host(eng.add_mod2pi(h0, ln)), P.add_lens(h0, 0.5))
    # batch broadcast of the addend
    hb = np.stack([h0, h0[::-1].copy()])
    out = eng.to_host(eng.add_mod2pi(hb, d))
    np.testing.assert_array_equal(out[1], (hb[1] + d) % (2 * np.pi))

    q = golden("quantize_96x128")
    m = q["mask"]
    for nm, h in (("rand", q["hologram"]), ("edge", q["hologram_edge"])):
        for ct in (256, 255, 200):
            np.testing.assert_array_equal(eng.to_host(eng.quantize(h, None, ct, _ffi.QUANT_ROUND_WRAP)), q[f"q1_{nm}_{ct}"])
            np.testing.assert_array_equal(eng.to_host(eng.quantize(h, m, ct, _ffi.QUANT_PIL_FLOAT)), q[f"q2_{nm}_{ct}"])
            np.testing.assert_array_equal(eng.to_host(eng.quantize(h, m, ct, _ffi.QUANT_FLOOR)), q[f"q3_{nm}_{ct}"])
            np.testing.assert_array_equal(eng.to_host(eng.quantize(h, None, ct, _ffi.QUANT_FLOOR)), q[f"q4_{nm}_{ct}"])
    np.testing.assert_array_equal(eng.to_host(eng.quantize_grey(q["png"], m, 200)), q["q2png_200"])
    np.testing.assert_array_equal(eng.to_host(eng.quantize(q["preview_in"], None, 256, _ffi.QUANT_PREVIEW)), q["preview_L"])
    eng.close()


def check_expected_outcome(make_engine, golden, precision):
    g = golden("preview_trap")
    h = g["hologram"]
    eng = make_engine(h.shape, precision, 1)
    out = eng.to_host(eng.expected_outcome(h, 255))
    tol = 1e-12 if precision == "fp64" else 1e-5
    assert np.abs(out - g["preview"]).max() <= tol * 255
    eng.close()


def check_random_phasor_guess(make_engine, golden, precision):
    """make_initial_guess "random"/"zeros": exp(1j*2*pi*u)[/100] evaluated on the device from the
    host-drawn MT19937 stream equals the reference's planes (golden) to rounding."""
    g = golden("initial_guess_24x40")
    shape = (64, 64)
    eng = make_engine(shape, precision, 1)
    tol = 4e-16 if precision == "fp64" else 2e-7
    for kind in ("random", "zeros"):
        u, div = hl.uniform_stream_guess(kind, shape, 42)
        x = eng.to_host(eng.random_phasor_guess(u, div))[0]
        ref = hl.host_initial_guess(kind, shape, 42)
        assert np.abs(x - ref).max() <= tol * np.abs(ref).max()
    # the stream itself is the reference's: first 24*40 draws reproduce the golden plane
    u, _ = hl.uniform_stream_guess("random", (24, 40), 42)
    np.testing.assert_array_equal(np.exp(1j * 2 * np.pi * u), g["random"])
    eng.close()


def check_gif_snapshots(make_engine_unused, precision, tmp_path, gs_fn, gd_fn, ns):
    """args.gif: the chunked run writes the reference's frames and returns the uninterrupted result."""
    import PIL.Image as im
    t = synthetic.shapes_target((128, 128))
    for which, (fn, kw) in enumerate(((gs_fn, {}), (gd_fn, dict(learning_rate=0.01)))):
        for gtype in ("i", "h"):
            d = tmp_path / f"alg{which}_{gtype}"
            d.mkdir()
            a = ns(max_loops=7, gif=True, gif_skip=3, gif_type=gtype, gif_source_dir=str(d), precision=precision, **kw)
            holo, exp, errs = fn(t, a)
            b = ns(max_loops=7, precision=precision, **kw)
            holo0, exp0, errs0 = fn(t, b)
            assert sorted(p.name for p in d.iterdir()) == ["0.png", "1.png", "2.png"]      # iterations 0, 3, 6
            assert len(errs) == 7
            tol = 1e-9 if precision == "fp64" else 2e-3
            assert np.max(np.abs(np.array(errs) - np.array(errs0)) / np.array(errs0)) < tol
            # the last frame is the final state (iteration 6)
            last = np.array(im.open(d / "2.png"))
            ref = (P.preview_to_L(exp) if gtype == "i" else P.preview_to_L((holo + np.pi) * 256 / (2 * np.pi)))
            assert np.mean(last != ref) < 1e-3
        # the loop condition holds across the chunk borders of a snapshot run (algorithms.py:29,83): with gif_skip = 1
        # every chunk is ONE iteration, and the run must still stop where the uninterrupted one does
        b = ns(max_loops=12, precision=precision, **kw)
        _, _, errs0 = fn(t, b)
        tol = float(0.5 * (errs0[4] + errs0[5]))
        stop = int(np.argmax(~(np.array(errs0) > tol))) + 1
        assert 1 < stop < 12
        for skip in (1, 2):
            d = tmp_path / f"alg{which}_tol{skip}"
            d.mkdir()
            a = ns(max_loops=12, tolerance=tol, gif=True, gif_skip=skip, gif_type="i", gif_source_dir=str(d), precision=precision, **kw)
            _, _, errs = fn(t, a)
            assert len(errs) == stop
            assert len(list(d.iterdir())) == (stop - 1) // skip + 1            # frames of iterations 0, skip, 2 skip, ... < stop
            if which == 1:
                c_ = ns(max_loops=12, tolerance=tol, precision=precision, **kw)
                fn(t, c_)
                assert a.learning_rate == c_.learning_rate


def check_single_trap_and_frames(make_engine, golden):
    g = golden("preview_trap")
    r, cc = [int(v) for v in g["trap_rc"]]
    eng = make_engine((64, 64), "fp64", 1)
    ph = eng.to_host(eng.single_trap_phase(r, cc, (192, 256)))
    assert circ(ph, g["trap_phase"]).max() < 1e-11
    assert np.all(ph > -np.pi - 1e-15) and np.all(ph <= np.pi + 1e-15)
    dots = synthetic.movie_frame_dots(6, rescale_parameter=13.0)
    frames = eng.to_host(eng.trap_frames(dots, 6, (768, 1024)))
    np.testing.assert_array_equal(frames, synthetic.movie_frames(6, rescale_parameter=13.0))
    with pytest.raises(IndexError):
        eng.trap_frames(np.array([[0, 800, 3]]), 1, (768, 1024))
    eng.close()


def check_device_mt19937(make_engine):
    """The device continuation of CPython's MT19937 equals random.random() draw for draw, and leaves the
    module-level generator in the same state."""
    import random
    eng = make_engine((64, 64), "fp64", 1)
    for seed, shape in ((42, (64, 64)), (7, (3, 1000)), (2**40 + 12345, (1, 311)), (42.0, (2, 312)), (0, (1, 1))):
        u = eng.to_host(eng.python_random_uniform(seed, shape))
        after = random.random()
        random.seed(seed)
        ref = np.array([random.random() for _ in range(int(np.prod(shape)))]).reshape(shape)
        np.testing.assert_array_equal(u, ref)
        assert after == random.random()
    eng.close()
