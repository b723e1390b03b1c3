"""N>1 path on CPU: two gloo ranks shard an optical-trap movie frame-wise (contiguous blocks, no
data-path collective), run the kernels' host emulation on their block, and gather on rank 0.
The result must equal the single-process run frame for frame."""
import os
import sys

import numpy as np
import pytest

from tests.conftest import free_port
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _frames(n, shape=(64, 128)):
    from spatial_light_modulator_module_b200 import synthetic
    return synthetic.movie_frames(n, rescale_parameter=11.0, shape=shape)


def _factory(shape, precision, batch):
    from tests.emu.emu_engine import EmuEngine
    return EmuEngine(shape, precision, batch)


def _worker(rank, world, port, n_frames, out_dir, gather):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from spatial_light_modulator_module_b200 import generate_hologram_sequence as ghs, shared_host
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        run = lambda: ghs.sequence_holograms(_frames(n_frames), 5, precision="fp64", batch=2, want_expected=True,   # noqa: E731
                                             engine_factory=_factory, gather=gather)
        first = run()
        kept = first[0].copy()
        holos, exps, errors, (lo, hi) = run()            # (the first result is still alive: rank 0 must not hand its memory out again)
        np.testing.assert_array_equal(first[0], kept)
        pool_while_alive = len(shared_host._POOL)
        del first
        third = run()                                    # ... and now it may
        np.testing.assert_array_equal(third[0], holos)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), holos=holos, exps=exps, lo=lo, hi=hi,
                 errors=np.array([np.asarray(e) for e in errors]), pool=np.array([pool_while_alive, len(shared_host._POOL)]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames,gather", [(5, True), (4, "host"), (5, "device")])
def test_two_rank_movie_matches_single_process(tmp_path, n_frames, gather):
    """gather: through host memory shared by the ranks of one node (the default there) or device to device."""
    from spatial_light_modulator_module_b200 import generate_hologram_sequence as ghs
    port = free_port()
    mp.spawn(_worker, args=(2, port, n_frames, str(tmp_path), gather), nprocs=2, join=True)
    ref_h, ref_e, ref_err, (lo, hi) = ghs.sequence_holograms(_frames(n_frames), 5, precision="fp64", batch=3,
                                                            want_expected=True, engine_factory=_factory)
    assert (lo, hi) == (0, n_frames)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert (int(r0["lo"]), int(r0["hi"])) == (0, n_frames)           # rank 0 holds the gathered movie
    np.testing.assert_array_equal(r0["holos"], ref_h)
    np.testing.assert_array_equal(r0["exps"], ref_e)
    np.testing.assert_array_equal(r0["errors"], np.array(ref_err))
    lo1, hi1 = int(r1["lo"]), int(r1["hi"])                           # rank 1 computed its own block ...
    assert (lo1, hi1) == ((n_frames + 1) // 2, n_frames)
    assert r1["holos"].shape[0] == 0                                  # ... and holds none of the movie
    # shared host memory: two segments per result set (holograms, expected); two sets while the first result lived, and no more after
    assert list(r0["pool"]) == ([0, 0] if gather == "device" else [4, 4])
    np.testing.assert_array_equal(r1["errors"], np.array(ref_err)[lo1:hi1])


def _worker_frames(rank, world, port, n_frames, out_dir):
    """uint8 SLM frames (mask add + floor quantisation), rasterised trap targets, a batch callback."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from spatial_light_modulator_module_b200 import generate_hologram_sequence as ghs, synthetic
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        shape = (64, 128)
        mask = synthetic.random_mask(shape, seed=4)
        dots = synthetic.movie_frame_dots(n_frames, rescale_parameter=11.0, shape=shape)
        seen = []
        frames, _, errors, (lo, hi) = ghs.sequence_holograms(
            None, 4, precision="fp64", batch=2, engine_factory=_factory, output="uint8", mask=mask, ct2pi=200,
            trap_dots=(dots, n_frames, shape), on_batch=lambda a, b, h, e: seen.append((a, b, h.copy())))
        np.savez(os.path.join(out_dir, f"frames{rank}.npz"), frames=frames, lo=lo, hi=hi, seen_lo=np.array([s[0] for s in seen]),
                 seen_hi=np.array([s[1] for s in seen]))
        for a, b, h in seen:
            np.testing.assert_array_equal(h, frames[a:b])
        # without a callback the same frames are gathered through shared host memory: the same bytes on rank 0
        again, _, errors2, (lo2, hi2) = ghs.sequence_holograms(
            None, 4, precision="fp64", batch=2, engine_factory=_factory, output="uint8", mask=mask, ct2pi=200,
            trap_dots=(dots, n_frames, shape))
        assert (lo2, hi2) == (lo, hi) and len(errors2) == len(errors)
        if rank == 0:
            np.testing.assert_array_equal(again, frames)
    finally:
        dist.destroy_process_group()


def test_two_rank_uint8_frames_from_device_rasterised_targets(tmp_path):
    from oracle import numpy_port as P
    from spatial_light_modulator_module_b200 import generate_hologram_sequence as ghs, synthetic
    n_frames, shape = 5, (64, 128)
    port = free_port()
    mp.spawn(_worker_frames, args=(2, port, n_frames, str(tmp_path)), nprocs=2, join=True)
    holos, _, _, _ = ghs.sequence_holograms(_frames(n_frames), 4, precision="fp64", batch=3, engine_factory=_factory)
    mask = synthetic.random_mask(shape, seed=4)
    r0 = np.load(tmp_path / "frames0.npz")
    assert r0["frames"].dtype == np.uint8 and r0["frames"].shape == (n_frames,) + shape
    for i in range(n_frames):
        np.testing.assert_array_equal(r0["frames"][i], P.quantize_q3(holos[i], mask, 200))
    assert sorted(zip(r0["seen_lo"], r0["seen_hi"])) == [(0, 2), (2, 3), (3, 5)]     # every rank's batches reached the callback


def _shared_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from spatial_light_modulator_module_b200 import shared_host
    os.environ["SLM_SHARED_RESULT_BYTES"] = str(3 << 20)               # room for three 1 MiB segments
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        log = []
        keep = []
        for n in (1, 1, 2, 4, 1):                                      # MiB; results dropped at once except the second
            rows = 2 * n
            (arr,) = shared_host.shared_results(dist, [((rows, 1 << 19), np.uint8)], (rank * n, (rank + 1) * n), page_lock=False)
            arr[rank * n:(rank + 1) * n] = rank + 1 + 10 * len(log)     # every rank fills its rows ...
            dist.barrier()
            if rank == 0:                                              # ... and rank 0 sees all of them
                assert (arr[:n] == 1 + 10 * len(log)).all() and (arr[n:] == 2 + 10 * len(log)).all()
            if len(log) == 1:
                keep.append(arr)
            log.append((len(shared_host._POOL), sum(e["nbytes"] for e in shared_host._POOL) >> 20))
            del arr
            dist.barrier()
        np.save(os.path.join(out_dir, f"pool{rank}.npy"), np.array(log))
        names = [e["name"] for e in shared_host._POOL]
        np.save(os.path.join(out_dir, f"names{rank}.npy"), np.array(names))
    finally:
        dist.destroy_process_group()


def test_shared_result_pool(tmp_path):
    """shared_host: segments are used again once their array is gone, never while it lives, given back beyond the
    byte limit, and none is left in /dev/shm when the processes end."""
    mp.spawn(_shared_worker, args=(2, free_port(), str(tmp_path)), nprocs=2, join=True)
    p0, p1 = np.load(tmp_path / "pool0.npy"), np.load(tmp_path / "pool1.npy")
    # call 1: one 1 MiB segment; call 2 re-uses it (kept alive afterwards); call 3: + 2 MiB; call 4: + 4 MiB, over the limit, so the
    # free 2 MiB one goes (the 1 MiB one is held); call 5: a new 1 MiB segment beside the held one, and the free 4 MiB one goes
    assert [list(r) for r in p0] == [[1, 1], [1, 1], [2, 3], [2, 5], [2, 2]]
    np.testing.assert_array_equal(p0, p1)                              # the other rank maps and drops the same segments
    for name in np.load(tmp_path / "names0.npy"):
        assert not os.path.exists(os.path.join("/dev/shm", str(name).lstrip("/")))
