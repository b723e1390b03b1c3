"""Pin the CPU oracle (oracle/numpy_port.py) to outputs of the unmodified reference.

The fixtures in tests/golden were written by oracle/make_golden.py from the reference's own
functions.  The port uses the same numpy/scipy operations in the same order, so on the library
versions recorded in the fixtures the match is bit-exact; on other versions the last bits of
scipy's FFT may move and the dense-target GS runs (chaotic) are compared
only over their first iterations.
"""
import hashlib

import numpy as np
import pytest
import scipy

from oracle import numpy_port as P
from spatial_light_modulator_module_b200 import synthetic


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def same_libs(g):
    return str(g["numpy_version"]) == np.__version__ and str(g["scipy_version"]) == scipy.__version__


def check_run(g, holo, exp, errs, chaotic):
    if same_libs(g):
        np.testing.assert_array_equal(np.array(errs), g["errors"])
        np.testing.assert_array_equal(holo, g["hologram"])
        np.testing.assert_array_equal(exp, g["expected"])
    else:  # pragma: no cover - other library builds
        n = 4 if chaotic else len(errs)
        np.testing.assert_allclose(np.array(errs)[:n], g["errors"][:n], rtol=1e-9)


@pytest.mark.parametrize("name,chaotic", [
    ("gs_noise_128x128", True), ("gs_shapes_128x128", True), ("gs_traps_128x128", False),
    ("gs_noise_192x256", True), ("gs_shapes_192x256", True), ("gs_traps_192x256", False),
    ("gs_noise_float64_128x128", True), ("gs_noise_float32_128x128", True), ("gs_noise_uint16_128x128", True),
])
def test_gs_matches_reference(golden, name, chaotic):
    g = golden(name)
    holo, exp, errs = P.gs_run(g["target"], int(g["max_loops"]))
    assert len(errs) == int(g["max_loops"])
    check_run(g, holo, exp, errs, chaotic)


def test_gs_tolerance_stop(golden):
    g = golden("gs_shapes_tol_128x128")
    holo, exp, errs = P.gs_run(g["target"], int(g["max_loops"]), tolerance=float(g["tolerance"]))
    assert len(errs) == len(g["errors"]) < int(g["max_loops"])
    check_run(g, holo, exp, errs, False)


def test_gs_zero_loops_raises():
    with pytest.raises(UnboundLocalError):
        P.gs_run(synthetic.traps_target((128, 128)), 0)


def test_gs_dtype_chain():
    """uint8 target -> float16 amplitude -> complex64 first ifft2."""
    st = P.gs_setup(synthetic.noise_target((64, 64)))
    assert st.target_amp.dtype == np.float16
    assert st.A.dtype == np.complex64
    assert P.gs_first_phasor(st).dtype == np.complex128


GD_CASES = ["gd_noise_random_128x128", "gd_shapes_fourier_192x256", "gd_traps_unsettle_128x128",
            "gd_noise_wa2int_128x128", "gd_noise_wa05_128x128", "gd_shapes_old_128x128",
            "gd_shapes_unnormed_128x128", "gd_shapes_zeros_128x128", "gd_shapes_ones_128x128",
            "gd_traps_tol_128x128"]


def gd_kwargs(g):
    kw = {}
    for k in g.files:
        if k.startswith("arg_"):
            v = g[k][()]
            kw[k[4:]] = v.item() if hasattr(v, "item") else v
    loops = kw.pop("max_loops")
    kw.pop("tolerance", None)
    if "initial_guess" in kw:
        kw["initial_guess_kind"] = str(kw.pop("initial_guess"))
    return loops, kw


@pytest.mark.parametrize("name", GD_CASES)
def test_gd_matches_reference(golden, name):
    g = golden(name)
    loops, kw = gd_kwargs(g)
    with np.errstate(all="ignore"):
        tol = float(g["arg_tolerance"]) if "arg_tolerance" in g.files else 0
        holo, exp, errs, lr = P.gd_run(g["target"], loops, tolerance=tol, **kw)
    assert len(errs) == len(g["errors"])
    assert lr == float(g["final_learning_rate"])
    if same_libs(g):
        np.testing.assert_array_equal(np.array(errs), g["errors"])
        np.testing.assert_array_equal(holo, g["hologram"])
        np.testing.assert_array_equal(exp, g["expected"])
    else:  # pragma: no cover
        np.testing.assert_allclose(np.array(errs), g["errors"], rtol=1e-9)


def test_initial_guesses(golden):
    g = golden("initial_guess_24x40")
    t = g["target"]
    ones = np.ones(t.shape)
    for kind in ("random", "old", "unnormed", "zeros", "ones", "fourier"):
        np.testing.assert_array_equal(P.initial_guess(kind, ones, t, 42), g[kind])
    np.testing.assert_array_equal(P.initial_guess("random", ones, t, 7), g["random_seed7"])
    np.testing.assert_array_equal(P.initial_guess("random", ones, t, 42.0), g["random_seed_float"])
    np.testing.assert_array_equal(P.initial_guess("random", ones, t, 2**40 + 12345), g["random_seed_big"])
    with pytest.raises(ValueError):
        P.initial_guess("nope", ones, t, 1)


def test_full_size_curves(golden):
    g = golden("gs_noise_512x512_curves")
    holo, exp, errs = P.gs_run(synthetic.noise_target((512, 512), seed=0), 20)
    if same_libs(g):
        np.testing.assert_array_equal(np.array(errs), g["errors"])
        assert sha(holo) == str(g["hologram_sha"])
    else:  # pragma: no cover
        np.testing.assert_allclose(np.array(errs)[:4], g["errors"][:4], rtol=1e-9)


def test_benched_path_fixtures(golden):
    """Round-2 fixtures (oracle/make_golden.py --benched-path): config 2 at the metric shape, 100 iterations, and the
    8-bit frames of the reference's hologram (mask add + floor / PIL-float32 quantisation)."""
    g = golden("gd_noise_1024x1024_curves")
    holo, exp, errs, _ = P.gd_run(synthetic.noise_target((1024, 1024), seed=0), 100)
    mask = synthetic.random_mask((1024, 1024), seed=1)
    if same_libs(g):
        np.testing.assert_array_equal(np.array(errs), g["errors"])
        assert sha(holo) == str(g["hologram_sha"])
        assert sha(P.quantize_q3(holo, mask, 256)) == str(g["q3_sha"])
        np.testing.assert_array_equal(P.quantize_q2(holo, mask, 256)[::4, ::4], g["q2_sub"])
    else:  # pragma: no cover
        np.testing.assert_allclose(np.array(errs), g["errors"], rtol=1e-9)
    g = golden("gs_noise_1024x1024_curves")
    _, _, errs = P.gs_run(synthetic.noise_target((1024, 1024), seed=0), 10)
    if same_libs(g):
        np.testing.assert_array_equal(np.array(errs), g["errors"])
    else:  # pragma: no cover
        np.testing.assert_allclose(np.array(errs)[:4], g["errors"][:4], rtol=1e-9)


def test_analytic(golden):
    g = golden("analytic")
    d = P.deflect_phase((1.0, 2.0))
    assert sha(d) == str(g["deflect_sha"])
    assert sha(P.deflect_phase((-0.5, 0.25))) == str(g["deflect2_sha"])
    ln = P.lens_phase(0.5, (768, 1024))
    assert ln.dtype == np.uint8 and set(np.unique(ln)) <= set(range(7))
    assert sha(ln) == str(g["lens_sha"])
    np.testing.assert_array_equal(P.lens_phase(-1.25, (96, 128)), g["lens2"])
    h0 = np.random.default_rng(int(g["h0_seed"])).uniform(-np.pi, np.pi, size=(768, 1024))
    assert sha(P.deflect_hologram(h0, (1.0, 2.0))) == str(g["deflected_sha"])
    assert sha(P.add_lens(h0, 0.5)) == str(g["lensed_sha"])


def test_quantisers(golden):
    g = golden("quantize_96x128")
    m = g["mask"]
    for nm, h in (("rand", g["hologram"]), ("edge", g["hologram_edge"])):
        for ct in (256, 255, 200):
            np.testing.assert_array_equal(P.quantize_q1(h, ct), g[f"q1_{nm}_{ct}"])
            np.testing.assert_array_equal(P.quantize_q2(h, m, ct), g[f"q2_{nm}_{ct}"])
            np.testing.assert_array_equal(P.quantize_q3(h, m, ct), g[f"q3_{nm}_{ct}"])
            np.testing.assert_array_equal(P.quantize_q3(h, None, ct), g[f"q4_{nm}_{ct}"])
    np.testing.assert_array_equal(P.quantize_q2_png(g["png"], m, 200), g["q2png_200"])
    np.testing.assert_array_equal(P.preview_to_L(g["preview_in"]), g["preview_L"])


def test_preview_and_trap(golden):
    g = golden("preview_trap")
    np.testing.assert_array_equal(P.expected_outcome_preview(g["hologram"], 255), g["preview"])
    r, c = g["trap_rc"]
    np.testing.assert_array_equal(P.single_trap_phase((192, 256), int(r), int(c)), g["trap_phase"])


def test_movie_frames_match_port():
    f = synthetic.movie_frames(5, rescale_parameter=7.0)
    for k in range(5):
        np.testing.assert_array_equal(f[k], P.traps_frame(P.two_circulating_dots(7.0 * k)))
