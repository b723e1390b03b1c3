"""Optical-trap movies: a sequence of target frames -> a sequence of GS holograms
(reference: ``generate_hologram_sequence.py:10-32``).

Frames are independent (the reference's loop carries no state from frame to frame), so they are
processed in device batches and, under ``torch.distributed``, split into contiguous blocks over
the ranks (one process per GPU) with no collective on the data path; only the finished holograms
are gathered.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Tuple

import numpy as np

from . import host_logic as hl


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def sequence_holograms(frames, max_loops: int, tolerance: float = 0.0, precision: str = "fp32",
                       batch: int = 32, want_expected: bool = False, engine_factory: Optional[Callable] = None,
                       gather: bool = True, inc_amp=None, warm_start: bool = False):
    """GS holograms of ``frames`` (uint8 [F,H,W]); ``inc_amp``: illumination amplitude plane [H,W] shared by all
    frames (algorithms.py:14-19), None = uniform.

    ``warm_start`` (an extension, off by default because it changes the results): every frame of a rank's block
    starts from the previous frame's hologram (``B = inc * exp(1j * hologram)``) instead of the reference's
    ``ifft2(sqrt(target))`` setup, which lets a slowly moving trap pattern converge in a few iterations; the frames
    are then processed one after the other.

    Under an initialised process group every rank passes the SAME ``frames`` and computes only its
    block; with ``gather`` the full results are returned on rank 0 (other ranks get their own
    block).  Returns ``(holograms [n,H,W] float64, expected or None, errors list, (lo, hi))``.
    """
    frames = np.asarray(frames)
    if frames.ndim != 3:
        raise ValueError("frames must be [F,H,W]")
    dist = _dist()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    lo, hi = hl.shard_range(frames.shape[0], rank, world)
    n_local = hi - lo
    shape = frames.shape[1:]
    batch = max(1, min(batch, max(n_local, 1)))
    if engine_factory is None:
        from .engine import get_engine
        eng = get_engine(shape, precision, batch)
    else:
        eng = engine_factory(shape, precision, batch)
    holos = eng.host_empty((n_local,) + shape, np.float64)
    exps = eng.host_empty((n_local,) + shape, np.float64) if want_expected else None
    errors: List[np.ndarray] = []
    if warm_start:
        phasor = None
        for s in range(n_local):
            res = eng.gs(frames[lo + s:lo + s + 1], max_loops, tolerance, inc_amp=inc_amp, phasor0=phasor, want_expected=want_expected)
            holos[s] = eng.to_host(res.hologram)[0]
            if want_expected:
                exps[s] = eng.to_host(res.expected)[0]
            errors.extend(res.errors)
            phasor = eng.phase_phasor(res.hologram, inc_amp)
        n_local = 0                                      # nothing left for the batched loop below
    # the read-back of one batch (float64: 8 bytes per pixel and frame) runs beside the iterations of the next one
    pending = []
    for s in range(0, n_local, batch):
        e = min(s + batch, n_local)
        res = eng.gs(frames[lo + s:lo + e], max_loops, tolerance, inc_amp=inc_amp, want_expected=want_expected)
        for job in pending:
            job.join()
        pending = [eng.to_host_into(res.hologram, holos[s:e])]
        if want_expected:
            pending.append(eng.to_host_into(res.expected, exps[s:e]))
        errors.extend(res.errors)
    for job in pending:
        job.join()
    if dist and world > 1 and gather:
        holos, exps, errors = _gather_to_root(dist, frames.shape[0], shape, holos, exps, errors, max_loops)
        if rank == 0:
            lo, hi = 0, frames.shape[0]
    return holos, exps, errors, (lo, hi)


def _gather_to_root(dist, n_total, shape, holos, exps, errors, max_loops):
    """Final gather of the per-rank blocks (contiguous, rank order) onto rank 0."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    counts = [hl.shard_range(n_total, r, world) for r in range(world)]
    cap = max(h - l for l, h in counts)

    def gather_block(local: np.ndarray, tail_shape) -> Optional[np.ndarray]:
        pad = torch.zeros((cap,) + tuple(tail_shape), dtype=torch.float64, device=dev)
        if local.shape[0]:
            pad[:local.shape[0]] = torch.from_numpy(local).to(dev)
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, bufs, dst=0)
        if rank != 0:
            return None
        return np.concatenate([bufs[r][:h - l].cpu().numpy() for r, (l, h) in enumerate(counts)], axis=0)

    curves = np.full((len(errors), max_loops + 1), np.nan)
    for i, e in enumerate(errors):
        curves[i, 0] = len(e)
        curves[i, 1:1 + len(e)] = e
    all_h = gather_block(holos, shape)
    all_e = gather_block(exps, shape) if exps is not None else None
    all_c = gather_block(curves, (max_loops + 1,))
    if rank != 0:
        return holos, exps, errors
    return all_h, all_e, [row[1:1 + int(row[0])].copy() for row in all_c]


def generate_hologram_sequence(args):
    """Drop-in for the reference driver (generate_hologram_sequence.py:10-32): reads
    ``images/moving_traps/<source_dir>/<i>.png``, writes ``<i>.npy`` holograms (and preview PNGs).
    Optional ``args.batch`` / ``args.precision`` tune the engine."""
    from PIL import Image as im
    dest_dir_holograms = f"holograms/{args.source_dir}_{args.version}_holograms"
    dest_dir_preview = f"images/moving_traps/{args.source_dir}_{args.version}_preview"
    for dest_dir in [dest_dir_holograms, dest_dir_preview]:
        if not os.path.exists(dest_dir):
            os.makedirs(dest_dir, exist_ok=True)
    source_dir_path = f"images/moving_traps/{args.source_dir}"
    files = os.listdir(source_dir_path)
    frames = np.stack([np.array(im.open(f"{source_dir_path}/{i}.png")) for i in range(len(files))])
    if frames.dtype != np.uint8 or frames.ndim != 3:
        raise ValueError("trap frames must be single-channel 8-bit images")
    inc = None
    if getattr(args, "incomming_intensity", "uniform") != "uniform":
        from .algorithms import _illumination
        inc = _illumination(args, frames.shape[1:])
    holos, exps, errors, (lo, hi) = sequence_holograms(
        frames, int(args.max_loops), float(args.tolerance), getattr(args, "precision", None) or
        os.environ.get("SLM_PRECISION", "fp32"), int(getattr(args, "batch", 32)), bool(args.preview), gather=False,
        inc_amp=inc)
    from .display_holograms import preview_to_grey
    for k in range(hi - lo):
        i = lo + k
        print(f"\rcreating {i}. hologram ", end="")
        np.save(f"{dest_dir_holograms}/{i}.npy", holos[k])
        if args.preview:
            im.fromarray(preview_to_grey(exps[k])).save(f"{dest_dir_preview}/{i}.png")
    plot_error_evolution([list(e) for e in errors])
    return errors


def plot_error_evolution(err_evl_list):
    """reference: generate_hologram_sequence.py:35-40 (skipped silently when matplotlib is absent)."""
    try:
        import matplotlib.pyplot as plt
    except Exception:
        return
    for i, err_evl in enumerate(err_evl_list):
        plt.plot(err_evl, label=i)
    plt.legend()
    plt.show()
