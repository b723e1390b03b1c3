"""Seeded synthetic targets for tests and benchmarks.

There is no network and the reference ships no images, so every workload uses targets of the
named shapes built here.  Pure numpy; no device code.
"""
from __future__ import annotations

import numpy as np

from . import constants as c


def noise_target(shape, seed=0) -> np.ndarray:
    """Dense uniform-noise uint8 target."""
    rng = np.random.default_rng(seed)
    return (rng.random(shape) * 255).astype(np.uint8)


def shapes_target(shape) -> np.ndarray:
    """Disc + bar on black, scaled to ``shape``."""
    h, w = shape
    t = np.zeros(shape, dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    t[(yy - h // 2) ** 2 + (xx - w // 3) ** 2 < (h // 6) ** 2] = 255
    t[h // 4:h // 4 + max(1, h // 25), w // 2:w // 2 + max(1, w // 5)] = 128
    return t


def traps_target(shape, points=None) -> np.ndarray:
    """Sparse optical-trap target: single white pixels (traps_images.py:87-91 semantics)."""
    h, w = shape
    t = np.zeros(shape, dtype=np.uint8)
    if points is None:
        points = [(int(w * 0.29), int(h * 0.26)), (int(w * 0.68), int(h * 0.65))]
    for (x, y) in points:
        t[round(y), round(x)] = 255
    return t


def two_circulating_dots(n, w=c.slm_width, h=c.slm_height):
    """Trap positions of movie frame parameter ``n`` -- generate_traps_image_sequence.py:48-58
    (with the reference's non-existent c.w/c.h read as slm_width/slm_height)."""
    t = n * (2 * np.pi) / 360
    return [
        (w * (1 / 2 + 1 / 3 * np.cos(t)), h * (1 / 2 + 1 / 3 * np.sin(t))),
        (w * (1 / 2 + 1 / 3 * np.cos(t + np.pi / 2)), h * (1 / 2 + 1 / 3 * np.sin(t + np.pi / 2))),
    ]


def circulating_dot_quarterized(n, w=c.slm_width, h=c.slm_height):
    """generate_traps_image_sequence.py:40-45."""
    t = n * (2 * np.pi) / 360
    w_, h_ = w // 2, h // 2
    return [(w_ * (1 / 2 + 1 / 10 * np.cos(t)), h_ * (1 / 2 + 1 / 10 * np.sin(t)))]


def two_circulating_dots_quarterized(n, w=c.slm_width, h=c.slm_height):
    """generate_traps_image_sequence.py:26-38."""
    t = n * (2 * np.pi) / 360
    w_, h_ = w // 2, h // 2
    return [
        (w_ * (1 / 2 + 1 / 3 * np.cos(t)), h_ * (1 / 2 + 1 / 3 * np.sin(t))),
        (w_ * (1 / 2 + 1 / 3 * np.cos(t + np.pi / 2)), h_ * (1 / 2 + 1 / 3 * np.sin(t + np.pi / 2))),
    ]


def movie_frames(number_of_frames, rescale_parameter=360 / 1024, parametrization=two_circulating_dots,
                 shape=(c.slm_height, c.slm_width), first=0) -> np.ndarray:
    """uint8 [F,H,W] stack: frame i = dots at parametrization(rescale*i)
    (generate_traps_image_sequence.py:61-71 + traps_images.py:24-29, minus the PNG round trip)."""
    h, w = shape
    out = np.zeros((number_of_frames, h, w), dtype=np.uint8)
    for k in range(number_of_frames):
        for (x, y) in parametrization(rescale_parameter * (first + k), w, h):
            out[k, round(y), round(x)] = 255
    return out


def random_mask(shape, seed=1) -> np.ndarray:
    """Stand-in wavefront-correction mask: uniform [0, 2pi) float64 (passes the range check at
    display_holograms.py:237)."""
    return np.random.default_rng(seed).uniform(0, 2 * np.pi, size=shape)


def movie_frame_dots(number_of_frames, rescale_parameter=360 / 1024, parametrization=two_circulating_dots,
                     shape=(c.slm_height, c.slm_width), first=0) -> np.ndarray:
    """(frame, y, x) of every white pixel of :func:`movie_frames`, for rasterising the movie directly on
    the device (Engine.trap_frames) instead of going through PNG files."""
    h, w = shape
    dots = []
    for k in range(number_of_frames):
        for (x, y) in parametrization(rescale_parameter * (first + k), w, h):
            dots.append((k, round(y), round(x)))
    return np.array(dots, dtype=np.int32).reshape(-1, 3)
