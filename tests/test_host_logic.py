"""Host-side logic of the drop-in shims against numpy / the oracle / the reference goldens."""
import random

import os

import numpy as np
import pytest

from oracle import numpy_port as P
from spatial_light_modulator_module_b200 import host_logic as hl


def test_amplitude_lut_is_numpys_float16_sqrt():
    lut = hl.amplitude_lut()
    t = np.arange(256, dtype=np.uint8)
    assert np.sqrt(t).dtype == np.float16
    np.testing.assert_array_equal(lut, np.abs(np.sqrt(t)).astype(np.float64))
    assert lut[255] != np.sqrt(255.0)       # rounded to float16


@pytest.mark.parametrize("wa", [1, 2, 0.5, np.float64(3.0)])
def test_gd_mask_lut_follows_numpy_promotion(wa):
    t = np.arange(256, dtype=np.uint8).reshape(16, 16)
    np.testing.assert_array_equal(hl.gd_mask_lut(wa)[t], 1 + wa * t / 255)


def test_classify_target():
    rng = np.random.default_rng(0)
    t8 = (rng.random((8, 8)) * 255).astype(np.uint8)
    assert hl.classify_target(t8)[0] == "u8"
    for dt, c64 in ((np.float64, False), (np.float32, True), (np.uint16, True), (np.int32, False)):
        kind, treal, amp, flag = hl.classify_target(t8.astype(dt))
        st = P.gs_setup(t8.astype(dt))
        assert kind == "real" and flag == c64 == (st.A.dtype == np.complex64)
        np.testing.assert_array_equal(amp, np.abs(st.target_amp).astype(np.float64))


@pytest.mark.parametrize("lr,unsettle,loops", [(0.005, 0, 10), (0.01, 2, 12), (0.01, 3, 10), (0.5, 1, 5), (0.1, 4, 30)])
def test_learning_rate_schedule_matches_oracle(lr, unsettle, loops):
    during, after = hl.learning_rate_schedule(lr, unsettle, loops)
    st = P.gd_setup(np.ones((4, 4), np.uint8), learning_rate=lr, unsettle=unsettle, max_loops=loops, x0=np.ones((4, 4)))
    for k in range(loops):
        assert during[k] == st.learning_rate
        st.i += 1
        if st.unsettle and st.i % int(round(st.max_loops / (st.unsettle + 1))) == 0:
            st.learning_rate *= 2
        assert after[k + 1] == st.learning_rate


def test_learning_rate_schedule_zero_period():
    with pytest.raises(ZeroDivisionError):
        hl.learning_rate_schedule(0.1, 5, 1)


def test_initial_guesses_match_reference(golden):
    g = golden("initial_guess_24x40")
    shape = g["target"].shape
    for kind in ("random", "old", "unnormed", "zeros", "ones"):
        np.testing.assert_array_equal(hl.host_initial_guess(kind, shape, 42), g[kind])
    np.testing.assert_array_equal(hl.host_initial_guess("random", shape, 7), g["random_seed7"])
    np.testing.assert_array_equal(hl.host_initial_guess("random", shape, 42.0), g["random_seed_float"])
    np.testing.assert_array_equal(hl.host_initial_guess("random", shape, 2**40 + 12345), g["random_seed_big"])
    assert hl.host_initial_guess("fourier", shape, 1) is None
    with pytest.raises(ValueError, match="unknown type of initial guess"):
        hl.host_initial_guess("nope", shape, 1)


def test_random_stream_leaves_module_state_like_the_reference_loop():
    u = hl.python_random_stream(123, 1000)
    after = random.random()
    random.seed(123)
    ref = [random.random() for _ in range(1000)]
    np.testing.assert_array_equal(u, ref)
    assert after == random.random()


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            blocks = [hl.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [h - l for l, h in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_scalars():
    from spatial_light_modulator_module_b200 import constants as c
    const, sy, sx = hl.deflect_scalars((1.0, 2.0), c.px_distance, c.wavelength, c.u)
    assert const == 2 * np.pi * c.px_distance / c.wavelength
    assert sy == np.sin(2.0 * c.u) and sx == np.sin(1.0 * c.u)
    k, f2 = hl.lens_scalars(0.5, c.wavelength)
    assert k == 2 * np.pi * 0.5 / c.wavelength and f2 == 0.25


def test_snapshot_chunks():
    assert hl.snapshot_chunks(7, 3) == [1, 3, 3]
    assert hl.snapshot_chunks(10, 1) == [1] * 10
    assert hl.snapshot_chunks(5, 10) == [1, 4]
    assert hl.snapshot_chunks(1, 4) == [1]
    for loops, skip in ((20, 6), (9, 2), (3, 7)):
        ch = hl.snapshot_chunks(loops, skip)
        assert sum(ch) == loops
        ends = np.cumsum(ch) - 1
        assert [e for e in ends if e % skip == 0] == [i for i in range(loops) if i % skip == 0]
    with pytest.raises(ZeroDivisionError):
        hl.snapshot_chunks(5, 0)


def test_command_lines_match_the_reference_defaults():
    """generate_hologram.py:222-371 and generate_hologram_sequence.py:43-107: flags, defaults and the attributes the
    drivers add after parsing."""
    from spatial_light_modulator_module_b200 import generate_hologram as gh, generate_hologram_sequence as ghs
    a = gh.build_parser().parse_args([])
    assert (a.img_name, a.incomming_intensity, a.initial_guess, a.destination_directory) == (None, "uniform", "random", "holograms")
    assert (a.algorithm, a.tolerance, a.max_loops, a.learning_rate, a.white_attention, a.unsettle) == ("gerchberg_saxton", 0, 42, 0.005, 1, 0)
    assert (a.gif, a.gif_type, a.gif_skip, a.plot_error, a.preview, a.deflect, a.lens) == (False, "i", 1, False, False, None, None)
    assert isinstance(a.white_attention, int)            # the int default wraps in uint8; "-wa 1" gives a float
    b = gh.build_parser().parse_args(["duck.png", "-alg", "gradient_descent", "-l", "100", "-wa", "1", "-deflect", "1", "2", "-lens", "0.5", "-q", "-i"])
    assert (b.img_name, b.algorithm, b.max_loops, b.deflect, b.lens, b.quarterize, b.invert) == ("duck.png", "gradient_descent", 100, [1.0, 2.0], 0.5, True, True)
    assert isinstance(b.white_attention, float)
    s = ghs.build_parser().parse_args(["seq", "-ct2pi", "256"])
    assert (s.source_dir, s.version, s.incomming_intensity, s.correspond_to2pi, s.tolerance, s.max_loops, s.preview) == \
        ("seq", None, "uniform", 256, 0, 5, False)
    with pytest.raises(SystemExit):
        ghs.build_parser().parse_args(["seq"])            # -ct2pi is required


def test_gif_directories_and_assembly(tmp_path, monkeypatch):
    """generate_hologram.py:90-99,206-219: where the frames go, and the GIF made of them."""
    import argparse
    from PIL import Image as im
    from spatial_light_modulator_module_b200 import generate_hologram as gh
    monkeypatch.chdir(tmp_path)
    a = argparse.Namespace(gif_type="i")
    gh.add_gif_dirs(a)
    assert (a.gif_dest_dir, a.gif_source_dir) == ("images", "images/gif_source") and os.path.isdir("images/gif_source")
    h = argparse.Namespace(gif_type="h")
    gh.add_gif_dirs(h)
    assert h.gif_source_dir == "holograms/gif_source"
    for i in range(3):
        im.fromarray(np.full((8, 8), 60 * i, dtype=np.uint8)).save(f"{a.gif_source_dir}/{i}.png")
    gh.create_gif(a.gif_source_dir, "images/out.gif")
    g = im.open("images/out.gif")
    assert getattr(g, "n_frames", 1) == 3
    gh.remove_files_in_dir(a.gif_source_dir)
    assert os.listdir(a.gif_source_dir) == []
