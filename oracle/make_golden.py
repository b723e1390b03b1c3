"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Usage:  python -m oracle.make_golden

The reference ships no golden vectors, so the fixtures are outputs of the
reference's own functions on seeded synthetic inputs.  Library versions are stored in every
file because scipy's FFT backend and numpy's ufunc loops define the last bits.
Large planes are stored as a strided subsample plus a SHA-256 of the full array.
"""
from __future__ import annotations

import argparse
import contextlib
import hashlib
import io
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_loader  # noqa: E402
from spatial_light_modulator_module_b200 import synthetic  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def versions():
    import PIL
    import scipy
    return dict(numpy_version=np.__version__, scipy_version=scipy.__version__, pillow_version=PIL.__version__)


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ns(**kw):
    base = dict(incomming_intensity="uniform", tolerance=0, max_loops=10, gif=False, gif_skip=1,
                gif_type="i", print_info=False, plot_error=False, initial_guess="random",
                random_seed=42, white_attention=1, learning_rate=0.005, unsettle=0,
                correspond_to2pi=256)
    base.update(kw)
    return argparse.Namespace(**base)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def save(name, **arrays):
    arrays.update({k: np.array(v) for k, v in versions().items()})
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def targets(shape):
    return {"noise": synthetic.noise_target(shape, seed=3), "shapes": synthetic.shapes_target(shape),
            "traps": synthetic.traps_target(shape)}


def benched_path_fixtures(alg):
    """Round 2: the exact configuration bench.py times (config 2 at the metric shape 1024x1024, 100 iterations)
    and the 8-bit frames of config 2 (mask add + quantise, move_traps.py:135-140 / display_holograms.py:253-266)
    from the reference's own holograms.  Frames are stored as a 4x4-strided subsample + SHA-256."""
    from PIL import Image as im
    sub = (slice(None, None, 16), slice(None, None, 16))
    sub4 = (slice(None, None, 4), slice(None, None, 4))
    for shape in ((1024, 1024), (768, 1024)):
        t = synthetic.noise_target(shape, seed=0)
        mask = synthetic.random_mask(shape, seed=1)
        holo, exp, errs = quiet(alg.gradient_descent, t, ns(max_loops=100))
        q3 = ((holo + mask) % (2 * np.pi) * 256 / (2 * np.pi)).astype(np.uint8)            # move_traps.py:135-140
        q2 = np.array(im.fromarray((holo + mask) % (2 * np.pi) / (2 * np.pi) * 256).convert("L"))   # display_holograms.py:257-260
        extra = dict(errors=np.array(errs), hologram_sub=holo[sub], expected_sub=exp[sub], hologram_sha=np.array(sha(holo)),
                     seed=np.array(0)) if shape == (1024, 1024) else {}
        save(f"gd_noise_{shape[0]}x{shape[1]}_{'curves' if extra else 'frames'}", q3_sub=q3[sub4], q2_sub=q2[sub4],
             q3_sha=np.array(sha(q3)), mask_seed=np.array(1), ct2pi=np.array(256), **extra)
    t = synthetic.noise_target((1024, 1024), seed=0)
    holo, exp, errs = quiet(alg.gerchberg_saxton, t, ns(max_loops=10))
    save("gs_noise_1024x1024_curves", errors=np.array(errs), hologram_sub=holo[sub], expected_sub=exp[sub],
         hologram_sha=np.array(sha(holo)), seed=np.array(0))


def slab_fixtures(alg):
    """Round 2: one 8192 x 8192 plane (lines only the slab path holds) through the reference's GD and GS -- the error
    curves and a 64x64-strided subsample of hologram / expected.  Minutes of numpy time and ~10 GB of memory."""
    n = 8192
    sub = (slice(None, None, 64), slice(None, None, 64))
    t = synthetic.traps_target((n, n), [(1000, 2000), (6000, 5000), (4096, 700)])
    holo, exp, errs = quiet(alg.gradient_descent, t, ns(max_loops=5))
    save("gd_traps_8192_slab", errors=np.array(errs), hologram_sub=holo[sub], expected_sub=exp[sub], hologram_sha=np.array(sha(holo)))
    holo, exp, errs = quiet(alg.gerchberg_saxton, t, ns(max_loops=6))
    save("gs_traps_8192_slab", errors=np.array(errs), hologram_sub=holo[sub], expected_sub=exp[sub], hologram_sha=np.array(sha(holo)))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    alg = reference_loader.load("algorithms")
    if "--slab" in sys.argv:                     # only the 8192^2 fixtures of the slab path
        slab_fixtures(alg)
        return
    if "--benched-path" in sys.argv:             # only the round-2 fixtures (the others are unchanged)
        benched_path_fixtures(alg)
        return
    gh = reference_loader.load("generate_hologram")
    wfc = reference_loader.load("wavefront_correction")
    dh = reference_loader.load("display_holograms")

    # ---- GS, small planes stored whole --------------------------------------------------
    for shape, loops in (((128, 128), 12), ((192, 256), 8)):
        for kind, t in targets(shape).items():
            holo, exp, errs = quiet(alg.gerchberg_saxton, t, ns(max_loops=loops))
            save(f"gs_{kind}_{shape[0]}x{shape[1]}", target=t, hologram=holo, expected=exp,
                 errors=np.array(errs), max_loops=np.array(loops))
    # non-uint8 targets: dtype chain of np.sqrt
    t = synthetic.noise_target((128, 128), seed=5)
    for dt in (np.float64, np.float32, np.uint16):
        tt = t.astype(dt) * (257 if dt == np.uint16 else 1)
        holo, exp, errs = quiet(alg.gerchberg_saxton, tt, ns(max_loops=6))
        save(f"gs_noise_{np.dtype(dt).name}_128x128", target=tt, hologram=holo, expected=exp,
             errors=np.array(errs), max_loops=np.array(6))
    # tolerance stop
    t = synthetic.shapes_target((128, 128))
    holo, exp, errs = quiet(alg.gerchberg_saxton, t, ns(max_loops=40, tolerance=3900.0))
    save("gs_shapes_tol_128x128", target=t, hologram=holo, expected=exp, errors=np.array(errs),
         max_loops=np.array(40), tolerance=np.array(3900.0))

    # ---- GD, small planes ------------------------------------------------------------------
    gd_cases = {
        "gd_noise_random_128x128": ((128, 128), "noise", dict(max_loops=20)),
        "gd_shapes_fourier_192x256": ((192, 256), "shapes", dict(max_loops=12, initial_guess="fourier")),
        "gd_traps_unsettle_128x128": ((128, 128), "traps", dict(max_loops=12, unsettle=2, learning_rate=0.01)),
        "gd_noise_wa2int_128x128": ((128, 128), "noise", dict(max_loops=8, white_attention=2)),
        "gd_noise_wa05_128x128": ((128, 128), "noise", dict(max_loops=8, white_attention=0.5)),
        "gd_shapes_old_128x128": ((128, 128), "shapes", dict(max_loops=6, initial_guess="old")),
        "gd_shapes_unnormed_128x128": ((128, 128), "shapes", dict(max_loops=6, initial_guess="unnormed")),
        "gd_shapes_zeros_128x128": ((128, 128), "shapes", dict(max_loops=6, initial_guess="zeros")),
        "gd_shapes_ones_128x128": ((128, 128), "shapes", dict(max_loops=6, initial_guess="ones")),
        "gd_traps_tol_128x128": ((128, 128), "traps", dict(max_loops=30, tolerance=5.0, unsettle=2, learning_rate=0.01)),
    }
    for name, (shape, kind, kw) in gd_cases.items():
        t = targets(shape)[kind]
        a = ns(**kw)
        holo, exp, errs = quiet(alg.gradient_descent, t, a)
        save(name, target=t, hologram=holo, expected=exp, errors=np.array(errs),
             final_learning_rate=np.array(a.learning_rate),
             **{f"arg_{k}": np.array(v) for k, v in kw.items()})

    # ---- initial guesses (first values + hash) -------------------------------------------
    t = synthetic.shapes_target((24, 40))
    ig = {}
    for kind in ("random", "old", "unnormed", "zeros", "ones", "fourier"):
        ig[kind] = alg.make_initial_guess(kind, np.ones(t.shape), t, 42)
    ig["random_seed7"] = alg.make_initial_guess("random", np.ones(t.shape), t, 7)
    ig["random_seed_float"] = alg.make_initial_guess("random", np.ones(t.shape), t, 42.0)
    ig["random_seed_big"] = alg.make_initial_guess("random", np.ones(t.shape), t, 2**40 + 12345)
    save("initial_guess_24x40", target=t, **ig)

    # ---- full-size runs: curves + subsample + hash ------------------------------------------
    sub = (slice(None, None, 16), slice(None, None, 16))
    t = synthetic.traps_target((768, 1024), [(300, 200), (700, 500)])
    holo, exp, errs = quiet(alg.gerchberg_saxton, t, ns(max_loops=50))
    save("gs_traps_768x1024_curves", errors=np.array(errs), hologram_sub=holo[sub], expected_sub=exp[sub],
         hologram_sha=np.array(sha(holo)), trap_points=np.array([(300, 200), (700, 500)]))
    t = synthetic.noise_target((512, 512), seed=0)
    holo, exp, errs = quiet(alg.gerchberg_saxton, t, ns(max_loops=20))
    save("gs_noise_512x512_curves", errors=np.array(errs), hologram_sub=holo[sub], expected_sub=exp[sub],
         hologram_sha=np.array(sha(holo)), seed=np.array(0))
    t = synthetic.noise_target((768, 1024), seed=0)
    holo, exp, errs = quiet(alg.gradient_descent, t, ns(max_loops=100))
    save("gd_noise_768x1024_curves", errors=np.array(errs), hologram_sub=holo[sub], expected_sub=exp[sub],
         hologram_sha=np.array(sha(holo)), seed=np.array(0))

    benched_path_fixtures(alg)

    # ---- analytic holograms --------------------------------------------------------------
    d = wfc.deflect_2pi((1.0, 2.0))
    d2 = wfc.deflect_2pi((-0.5, 0.25))
    ln = gh.lens(0.5, (768, 1024))
    ln2 = gh.lens(-1.25, (96, 128))
    h0 = np.random.default_rng(11).uniform(-np.pi, np.pi, size=(768, 1024))
    save("analytic", deflect_sub=d[sub], deflect_sha=np.array(sha(d)), deflect2_sub=d2[sub],
         deflect2_sha=np.array(sha(d2)), lens_sub=ln[sub], lens_sha=np.array(sha(ln)), lens2=ln2,
         deflected_sha=np.array(sha(gh.deflect_hologram(h0, (1.0, 2.0)))),
         lensed_sha=np.array(sha(gh.add_lens(h0, 0.5))), h0_seed=np.array(11))

    # ---- quantisers ----------------------------------------------------------------------
    rng = np.random.default_rng(21)
    h = rng.uniform(-np.pi, np.pi, size=(96, 128))
    # exact grey-level boundaries and wrap points, to exercise rounding modes
    edge = (np.arange(96 * 128) % 512 - 256) * (2 * np.pi / 256)
    h_edge = edge.reshape(96, 128)
    m = rng.uniform(0, 2 * np.pi, size=(96, 128))
    q = {}
    with tempfile.TemporaryDirectory() as td:
        for nm, hh in (("rand", h), ("edge", h_edge)):
            p = os.path.join(td, nm + ".npy")
            np.save(p, hh)
            for ct in (256, 255, 200):
                q[f"q1_{nm}_{ct}"] = wfc.convert_2pi_hologram_to_int_hologram(hh, ct)
                q[f"q2_{nm}_{ct}"] = np.array(dh.mask_hologram(p, m, ct))
                q[f"q3_{nm}_{ct}"] = (hh + m) % (2 * np.pi) * ct / (2 * np.pi)
                q[f"q3_{nm}_{ct}"] = q[f"q3_{nm}_{ct}"].astype(np.uint8)       # move_traps.py:135-140
                q[f"q4_{nm}_{ct}"] = (hh % (2 * np.pi) * ct / (2 * np.pi)).astype(np.uint8)  # show_hologram.py:9-11
        from PIL import Image as im
        png = (rng.random((96, 128)) * 255).astype(np.uint8)
        pp = os.path.join(td, "g.png")
        im.fromarray(png).save(pp)
        q["q2png_200"] = np.array(dh.mask_hologram(pp, m, 200))
        q["png"] = png
    exp = rng.uniform(0, 300, size=(96, 128))
    from PIL import Image as im
    q["preview_L"] = np.array(im.fromarray(exp).convert("L"))   # generate_hologram_sequence.py:29
    save("quantize_96x128", hologram=h, hologram_edge=h_edge, mask=m, preview_in=exp, **q)

    # ---- target preparation (PIL pipeline, generate_hologram.py:45-67,102-110,166-175) ----------
    from PIL import Image as im
    src = (np.random.default_rng(31).random((200, 300)) * 255).astype(np.uint8)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "images"))
        im.fromarray(src).save(os.path.join(td, "images", "src.png"))
        os.chdir(td)
        try:
            outs = {}
            for inv in (False, True):
                for qz in (False, True):
                    a = argparse.Namespace(invert=inv, quarterize=qz)
                    outs[f"prep_inv{int(inv)}_q{int(qz)}"] = gh.prepare_target("src.png", a)
        finally:
            os.chdir(cwd)
    save("prepare_target", src=src, **{k: v[::8, ::8] for k, v in outs.items()},
         **{k + "_sha": np.array(sha(v)) for k, v in outs.items()})

    # ---- expected-outcome preview and analytic single trap (generate_hologram.py:24-29, move_traps.py:64-68)
    h = np.random.default_rng(41).uniform(-np.pi, np.pi, size=(128, 128))
    from scipy.fft import fft2, ifft2
    e = np.abs(fft2(np.exp(1j * h))) ** 2
    prev = e / np.amax(e) * 255
    img = np.zeros((192, 256), dtype=np.uint8)
    img[37][101] = 255
    save("preview_trap", hologram=h, preview=prev, trap_phase=np.angle(ifft2(img)), trap_rc=np.array([37, 101]))


if __name__ == "__main__":
    main()
