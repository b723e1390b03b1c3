"""N>1 path on CPU: two gloo ranks shard an optical-trap movie frame-wise (contiguous blocks, no
data-path collective), run the kernels' host emulation on their block, and gather on rank 0.
The result must equal the single-process run frame for frame."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _frames(n, shape=(64, 128)):
    from spatial_light_modulator_module_b200 import synthetic
    return synthetic.movie_frames(n, rescale_parameter=11.0, shape=shape)


def _factory(shape, precision, batch):
    from tests.emu.emu_engine import EmuEngine
    return EmuEngine(shape, precision, batch)


def _worker(rank, world, port, n_frames, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from spatial_light_modulator_module_b200 import generate_hologram_sequence as ghs
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        holos, exps, errors, (lo, hi) = ghs.sequence_holograms(_frames(n_frames), 5, precision="fp64", batch=2,
                                                               want_expected=True, engine_factory=_factory)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), holos=holos, exps=exps, lo=lo, hi=hi,
                 errors=np.array([np.asarray(e) for e in errors]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [5, 4])
def test_two_rank_movie_matches_single_process(tmp_path, n_frames):
    from spatial_light_modulator_module_b200 import generate_hologram_sequence as ghs
    port = 29500 + (os.getpid() % 2000) + n_frames
    mp.spawn(_worker, args=(2, port, n_frames, str(tmp_path)), nprocs=2, join=True)
    ref_h, ref_e, ref_err, (lo, hi) = ghs.sequence_holograms(_frames(n_frames), 5, precision="fp64", batch=3,
                                                            want_expected=True, engine_factory=_factory)
    assert (lo, hi) == (0, n_frames)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert (int(r0["lo"]), int(r0["hi"])) == (0, n_frames)           # rank 0 holds the gathered movie
    np.testing.assert_array_equal(r0["holos"], ref_h)
    np.testing.assert_array_equal(r0["exps"], ref_e)
    np.testing.assert_array_equal(r0["errors"], np.array(ref_err))
    lo1, hi1 = int(r1["lo"]), int(r1["hi"])                           # rank 1 keeps its own block
    assert (lo1, hi1) == ((n_frames + 1) // 2, n_frames)
    np.testing.assert_array_equal(r1["holos"], ref_h[lo1:hi1])
