"""CPU restatement (numpy/scipy) of the reference GS/GD hologram path.

TEST INFRASTRUCTURE ONLY -- never imported by the shipped package.

Every function cites the reference lines it restates (paths relative to
/root/reference).  The restatement is organised as explicit *state machines*
(setup / one step / finish) so tests can teacher-force single iterations, which the
reference's monolithic loops do not allow.  The
arithmetic -- operand order, dtype promotion, scaling -- is the reference's.

Parity status: PINNED against the reference itself (tests/golden/*.npz produced by
oracle/make_golden.py from the unmodified /root/reference code, see
tests/test_oracle_golden.py).  The reference ships no golden vectors of its own.
"""
from __future__ import annotations

import random as _pyrandom
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np
from scipy.fft import fft2, ifft2

TWO_PI = 2 * np.pi

# constants.py:5-10
SLM_WIDTH = 1024
SLM_HEIGHT = 768
WAVELENGTH = 5.32e-7
PX_DISTANCE = 3.6e-5
FIRST_DIFF_MAX = WAVELENGTH / PX_DISTANCE
UNIT_ANGLE = FIRST_DIFF_MAX / 4


# --------------------------------------------------------------------------------------
# shared pieces
# --------------------------------------------------------------------------------------
def illumination_amplitude(shape, incomming_intensity="uniform"):
    """algorithms.py:14-19 / :65-70 -- sqrt of the illumination plane (ones when 'uniform').

    A non-'uniform' value is an already-loaded array here (the reference opens it with PIL).
    """
    if isinstance(incomming_intensity, str) and incomming_intensity == "uniform":
        plane = np.ones(shape)
    else:
        plane = np.asarray(incomming_intensity)
    return np.sqrt(plane)


def mean_square_error(actual, correct, norm):
    """algorithms.py:161-162 (error_f)."""
    return np.sum((actual - correct) ** 2) / norm


def unit_phasor(z):
    """exp(1j*angle(z)) exactly as written at algorithms.py:30,33."""
    return np.exp(1j * np.angle(z))


# --------------------------------------------------------------------------------------
# Gerchberg-Saxton  (algorithms.py:10-49)
# --------------------------------------------------------------------------------------
@dataclass
class GSState:
    target: np.ndarray            # demanded_output as given (dtype preserved)
    inc_amp: np.ndarray           # incomming_amplitude           (:19)
    target_amp: np.ndarray        # demanded_output_amplitude     (:21)  float16 for uint8 targets
    space_norm: int               # w*l                           (:22)
    norm: object                  # np.amax(demanded_output)      (:23)
    A: np.ndarray                 # SLM-plane field entering the next iteration (:27,:34)
    expected: Optional[np.ndarray] = None
    errors: List[float] = field(default_factory=list)


def gs_setup(target, incomming_intensity="uniform") -> GSState:
    """algorithms.py:14-27."""
    target = np.asarray(target)
    inc_amp = illumination_amplitude(target.shape, incomming_intensity)
    w, l = target.shape
    target_amp = np.sqrt(target)
    A = ifft2(target_amp)
    return GSState(target, inc_amp, target_amp, w * l, np.amax(target), A)


def gs_step(st: GSState) -> Tuple[np.ndarray, np.ndarray, float]:
    """One pass of the loop body algorithms.py:30-38.  Returns (C, expected, error) and
    advances ``st.A`` to the next SLM-plane field."""
    B = st.inc_amp * unit_phasor(st.A)                       # :30
    C = fft2(B)                                              # :31
    D = np.abs(st.target_amp) * unit_phasor(C)               # :33
    st.A = ifft2(D)                                          # :34
    expected = np.abs(C) ** 2                                # :36
    expected *= st.norm / expected.max()                     # :37
    err = mean_square_error(expected, st.target, st.space_norm)  # :38
    st.expected = expected
    st.errors.append(err)
    return C, expected, err


def gs_first_phasor(st: GSState) -> np.ndarray:
    """B of iteration 0 (algorithms.py:27,30) -- carries the complex64 / float32 quirk."""
    return st.inc_amp * unit_phasor(st.A)


def gs_run(target, max_loops, tolerance=0, incomming_intensity="uniform"):
    """algorithms.py:24-49 without the prints / gif side effects."""
    st = gs_setup(target, incomming_intensity)
    error = tolerance + 1
    i = 0
    while error > tolerance and i < max_loops:
        _, _, error = gs_step(st)
        i += 1
    if st.expected is None:
        raise UnboundLocalError("expected_outcome")           # reference behaviour, :49
    return np.angle(st.A), st.expected, st.errors


# --------------------------------------------------------------------------------------
# initial guesses (algorithms.py:115-158)
# --------------------------------------------------------------------------------------
def python_random_stream(seed, count) -> np.ndarray:
    """``count`` successive random.random() values after random.seed(seed).

    Bit-identical to the reference's per-pixel calls (algorithms.py:117-150); obtained by
    transplanting CPython's MT19937 state into numpy's RandomState.
    """
    _pyrandom.seed(seed)
    state = _pyrandom.getstate()
    rs = np.random.RandomState()
    rs.set_state(("MT19937", np.array(state[1][:-1], dtype=np.uint32), state[1][-1]))
    return rs.random_sample(count)


def initial_guess(kind, inc_amp, target, seed) -> np.ndarray:
    """algorithms.py:115-158 (make_initial_guess), vectorised over the same draw order."""
    h, w = target.shape
    if kind == "random":                                      # :118-124
        u = python_random_stream(seed, h * w).reshape(h, w)
        return np.exp(1j * 2 * np.pi * u)
    if kind == "old":                                         # :125-134 (real drawn first)
        u = python_random_stream(seed, 2 * h * w).reshape(h, w, 2)
        return np.sqrt(u[..., 0]) + 1j * np.sqrt(u[..., 1])
    if kind == "unnormed":                                    # :135-144
        u = python_random_stream(seed, 2 * h * w).reshape(h, w, 2)
        return ((u[..., 0] + 1j * u[..., 1]) - 0.5) * 2
    if kind == "zeros":                                       # :145-151
        u = python_random_stream(seed, h * w).reshape(h, w)
        return np.exp(1j * 2 * np.pi * u) / 100
    if kind == "ones":                                        # :152-153
        _pyrandom.seed(seed)
        return np.ones(target.shape) + 1j * np.zeros(target.shape)
    if kind == "fourier":                                     # :154-157
        _pyrandom.seed(seed)
        return inc_amp * np.exp(1j * np.angle(ifft2(np.sqrt(target))))
    raise ValueError("unknown type of initial guess")          # :158


# --------------------------------------------------------------------------------------
# gradient descent (algorithms.py:60-112, 179-185)
# --------------------------------------------------------------------------------------
def tangent_gradient_rows(dEdF, x):
    """algorithms.py:179-185 (dEdX_complex) applied to whole planes (the reference maps it
    over rows at :90; the arithmetic is elementwise so the result is identical)."""
    rE, iE = dEdF.real, dEdF.imag
    rx, ix = x.real, x.imag
    ax = abs(x)
    re_res = rE * (1 / ax - rx**2 / ax**3) + iE * (-(rx * ix) / ax**3)
    im_res = rE * (-(rx * ix) / ax**3) + iE * (1 / ax - ix**2 / ax**3)
    return re_res + 1j * im_res


@dataclass
class GDState:
    target: np.ndarray
    inc_amp: np.ndarray
    space_norm: int
    norm: object
    x: np.ndarray                 # 'input' (:75)
    mask: np.ndarray              # 1 + white_attention*T/255 (:80)
    learning_rate: float
    unsettle: int
    max_loops: int
    i: int = 0
    output: Optional[np.ndarray] = None
    errors: List[float] = field(default_factory=list)


def gd_setup(target, *, initial_guess_kind="random", seed=42, white_attention=1,
             learning_rate=0.005, unsettle=0, max_loops=100, incomming_intensity="uniform",
             x0=None) -> GDState:
    """algorithms.py:65-80."""
    target = np.asarray(target)
    inc_amp = illumination_amplitude(target.shape, incomming_intensity)
    w, l = target.shape
    x = initial_guess(initial_guess_kind, inc_amp, target, seed) if x0 is None else np.array(x0, dtype=complex)
    mask = 1 + white_attention * target / 255                 # :80  (uint8 wrap when wa is an int)
    return GDState(target, inc_amp, w * l, np.amax(target), x, mask, learning_rate, unsettle, max_loops)


def gd_step(st: GDState) -> Tuple[np.ndarray, np.ndarray, float]:
    """Loop body algorithms.py:84-92 and the unsettle rule :102-104.
    Returns (med_output, output, error); updates st.x, st.learning_rate, st.i."""
    med_output = fft2(st.x / abs(st.x) * st.inc_amp)          # :84
    output_unnormed = abs(med_output) ** 2                    # :85
    output = output_unnormed * st.norm / np.amax(output_unnormed)   # :86
    dEdF = ifft2(st.mask * med_output * (output - st.target)) * st.inc_amp  # :87-89
    dEdX = tangent_gradient_rows(dEdF, st.x)                  # :90
    st.x = st.x - st.learning_rate * dEdX                     # :91 (in-place in the reference)
    err = mean_square_error(output, st.target, st.space_norm)  # :92
    st.errors.append(err)
    st.output = output
    st.i += 1                                                 # :102
    if st.unsettle and st.i % int(round(st.max_loops / (st.unsettle + 1))) == 0:   # :103
        st.learning_rate *= 2                                 # :104
    return med_output, output, err


def gd_run(target, max_loops, tolerance=0, **kw):
    """algorithms.py:78-112 without prints / gif.  Returns (hologram, output, errors, final_lr)."""
    st = gd_setup(target, max_loops=max_loops, **kw)
    error = tolerance + 1
    while error > tolerance and st.i < max_loops:
        _, _, error = gd_step(st)
    if st.output is None:
        raise UnboundLocalError("output")                     # reference behaviour, :110
    return np.angle(st.x), st.output, st.errors, st.learning_rate


def complex_to_real_phase(z, correspond_to2pi=256):
    """algorithms.py:175-176."""
    return (np.angle(z) + np.pi) / (2 * np.pi) * correspond_to2pi


# --------------------------------------------------------------------------------------
# analytic holograms
# --------------------------------------------------------------------------------------
def deflect_phase(angle, shape=(SLM_HEIGHT, SLM_WIDTH)) -> np.ndarray:
    """wavefront_correction.py:440-449 (deflect_2pi): blazed grating, fixed SLM shape there.
    Same IEEE operations per pixel as the reference's scalar double loop."""
    x_angle, y_angle = angle
    h, w = shape
    const = 2 * np.pi * PX_DISTANCE / WAVELENGTH
    sy = np.sin(y_angle * UNIT_ANGLE)
    sx = np.sin(x_angle * UNIT_ANGLE)
    i = np.arange(h, dtype=np.float64)[:, None]
    j = np.arange(w, dtype=np.float64)[None, :]
    return (const * (sy * i + sx * j)) % TWO_PI


def lens_phase(focal_length, shape, uint8_quirk=True) -> np.ndarray:
    """generate_hologram.py:189-203 (lens).  With ``uint8_quirk`` the phase is truncated into a
    uint8 array exactly as the reference does (:192,:202); otherwise the float64 phase."""
    h, w = shape
    i = np.arange(h, dtype=np.float64)[:, None]
    j = np.arange(w, dtype=np.float64)[None, :]
    r = PX_DISTANCE * np.sqrt((i - h / 2) ** 2 + (j - w / 2) ** 2)
    phase = 2 * np.pi * focal_length / WAVELENGTH * (1 - np.sqrt(1 + r**2 / focal_length**2))
    phase = phase % TWO_PI
    if uint8_quirk:
        return phase.astype(np.uint8)
    return phase


def add_mod_2pi(hologram, addend):
    """generate_hologram.py:181,186: (hologram + addend) % (2*pi)."""
    return (hologram + addend) % TWO_PI


def deflect_hologram(hologram, angle):
    """generate_hologram.py:178-182."""
    return add_mod_2pi(hologram, deflect_phase(angle))


def add_lens(hologram, focal_len, uint8_quirk=True):
    """generate_hologram.py:185-186."""
    return add_mod_2pi(hologram, lens_phase(focal_len, hologram.shape, uint8_quirk))


def expected_outcome_preview(hologram, norm=255):
    """generate_hologram.py:24-29 (show_expected_outcome, numeric part)."""
    field_ = fft2(np.exp(1j * hologram))
    intensity = np.abs(field_) ** 2
    return intensity / np.amax(intensity) * norm


def single_trap_phase(shape, row, col):
    """move_traps.py:64-68 (update_hologram): angle(ifft2(one-hot*255))."""
    img = np.zeros(shape, dtype=np.uint8)
    img[row][col] = 255
    return np.angle(ifft2(img))


# --------------------------------------------------------------------------------------
# 8-bit quantisers
# --------------------------------------------------------------------------------------
def quantize_q1(hologram, ct2pi):
    """wavefront_correction.py:458-459: np.round(h*ct2pi/2pi).astype(uint8)."""
    return np.round(hologram * ct2pi / (2 * np.pi)).astype(np.uint8)


def _pil_float_to_L(arr):
    """PIL fromarray(float64)->mode 'F' (float32) -> convert('L'): clamp to [0,255], truncate."""
    f = np.asarray(arr, dtype=np.float64).astype(np.float32)
    f = np.where(f <= 0, np.float32(0), np.where(f >= 255, np.float32(255), f))
    return f.astype(np.uint8)


def quantize_q2(hologram, mask, ct2pi):
    """display_holograms.py:253-266 (mask_hologram, .npy branch)."""
    corrected = (hologram + mask) % (2 * np.pi)
    corrected = corrected / (2 * np.pi) * ct2pi
    return _pil_float_to_L(corrected)


def quantize_q2_png(grey_u8, mask, ct2pi):
    """display_holograms.py:259-264 (mask_hologram, image branch)."""
    arr = np.asarray(grey_u8).astype(np.int16)
    corrected = (arr + (mask / (2 * np.pi) * ct2pi)) % ct2pi
    return _pil_float_to_L(corrected)


def quantize_q3(hologram, mask, ct2pi):
    """move_traps.py:135-140 (display_hologram); show_hologram.py:7-13 is the same with mask=None
    and default ct2pi=255 (Q4)."""
    if mask is not None:
        hologram = hologram + mask
    return (hologram % (2 * np.pi) * ct2pi / (2 * np.pi)).astype(np.uint8)


def preview_to_L(expected):
    """generate_hologram_sequence.py:29: PIL fromarray(float64).convert('L')."""
    return _pil_float_to_L(expected)


# --------------------------------------------------------------------------------------
# trap-movie targets (traps_images.py:10-16,87-91; generate_traps_image_sequence.py:48-58)
# --------------------------------------------------------------------------------------
def two_circulating_dots(n, w=SLM_WIDTH, h=SLM_HEIGHT):
    """generate_traps_image_sequence.py:48-58 with c.w->slm_width, c.h->slm_height
    (the reference names attributes that do not exist)."""
    t = n * (2 * np.pi) / 360
    return [
        (w * (1 / 2 + 1 / 3 * np.cos(t)), h * (1 / 2 + 1 / 3 * np.sin(t))),
        (w * (1 / 2 + 1 / 3 * np.cos(t + np.pi / 2)), h * (1 / 2 + 1 / 3 * np.sin(t + np.pi / 2))),
    ]


def traps_frame(points, w=SLM_WIDTH, h=SLM_HEIGHT) -> np.ndarray:
    """traps_images.py:10-16 + dot() :87-91: single white pixels at round()ed coordinates."""
    img = np.zeros((h, w), dtype=np.uint8)
    for (x, y) in points:
        img[round(y), round(x)] = 255
    return img
