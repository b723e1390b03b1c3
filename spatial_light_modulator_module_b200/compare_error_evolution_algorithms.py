"""Error-vs-iteration curves of gradient descent and Gerchberg-Saxton on the same targets -- the numeric part
of the reference's ``compare_error_evolution_algorithms.py:10-20`` (the matplotlib plot stays with the
caller), batched: BASELINE config 4 runs it over hundreds of large targets at once."""
from __future__ import annotations

import os
from typing import List, Tuple

import numpy as np

from . import host_logic as hl
from .engine import get_engine


def fill_unnecessary_args(args):
    """reference: compare_error_evolution_GD_params.py:76-85."""
    args.gif = False
    args.gif_type = None
    args.gif_dir = None
    args.deflect = None
    args.lens = None
    args.correspond_to2pi = 256
    args.incomming_intensity = "uniform"
    args.print_info = False
    args.tolerance = 0


def error_evolution_curves(targets, args, batch: int = 16, engine_factory=None) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """``targets``: uint8 [B,H,W].  For every target: gradient_descent(target, args) then
    gerchberg_saxton(target, args) exactly as compare_error_evolution_algorithms.py:16,20 calls them
    (same ``args``, so a learning rate doubled by ``unsettle`` during GD carries over as in the
    reference), returning (GD curves, GS curves).  Targets are processed ``batch`` at a time on the device."""
    targets = np.asarray(targets)
    if targets.ndim != 3 or targets.dtype != np.uint8:
        raise ValueError("targets must be uint8 [B,H,W]")
    precision = getattr(args, "precision", None) or os.environ.get("SLM_PRECISION", "fp32")
    shape = targets.shape[1:]
    if engine_factory is None:
        eng = get_engine(shape, precision, min(batch, len(targets)), getattr(args, "device", None))
    else:                                                 # (the CPU test-suite runs the kernels' host emulation)
        eng = engine_factory(shape, precision, min(batch, len(targets)))
    loops = int(args.max_loops)
    during, after = hl.learning_rate_schedule(args.learning_rate, args.unsettle, loops)
    gd_curves, gs_curves = [], []
    for s in range(0, len(targets), batch):
        # (Sending chunk k+1 ahead of chunk k's loops was tried: the staging copy on the host then sits in front of the
        #  launches and the call got 5 % slower -- 1.82 instead of 1.72 s for 256 x 2048^2; the copy is 2.5 % of the call.)
        chunk = targets[s:s + batch]
        # make_initial_guess reseeds on every call (algorithms.py:117), so every target starts from the same plane
        if args.initial_guess in ("random", "zeros"):
            u = eng.python_random_uniform(args.random_seed, shape)
            plane = eng.random_phasor_guess(u, 100.0 if args.initial_guess == "zeros" else 1.0)          # [1,H,W] on the device
            x0 = eng._mem_repeat(plane, len(chunk))
        elif args.initial_guess == "fourier":
            hl.host_initial_guess("fourier", shape, args.random_seed)                                    # reseeds, like the reference
            x0 = eng.fourier_guess(chunk)
        else:
            g = hl.host_initial_guess(args.initial_guess, shape, args.random_seed)
            x0 = np.broadcast_to(g.astype(eng.complex_dtype), (len(chunk),) + tuple(shape)).copy()
        res_gd, _ = eng.gd(chunk, x0, during, loops, float(args.tolerance), white_attention=args.white_attention,
                           want_expected=False)
        gd_curves += res_gd.errors
        res_gs = eng.gs(chunk, loops, float(args.tolerance), want_expected=False)
        gs_curves += res_gs.errors
    args.learning_rate = after[max(len(c) for c in gd_curves)] if gd_curves else args.learning_rate
    return gd_curves, gs_curves
