"""cProfile of the drop-in call (gradient_descent -> numpy -> grey frame): where the host time goes."""
import sys, argparse, contextlib, io, cProfile, pstats, time
sys.path.insert(0, '.')
import numpy as np, torch
from spatial_light_modulator_module_b200 import algorithms, display_holograms, synthetic
shape = (1024, 1024)
ns = argparse.Namespace(incomming_intensity="uniform", tolerance=0, max_loops=100, gif=False, print_info=False, plot_error=False,
                        initial_guess="random", random_seed=42, white_attention=1, learning_rate=0.005, unsettle=0, precision="fp32", device=0)
targets = [synthetic.noise_target(shape, seed=i) for i in range(24)]
mask = synthetic.random_mask(shape)
def once(t, fn):
    with contextlib.redirect_stdout(io.StringIO()):
        h, e, errs = fn(t, ns)
    ns.learning_rate = 0.005
    return display_holograms.hologram_to_grey(h, mask, 256)
for name, fn in (("gd", algorithms.gradient_descent), ("gs", algorithms.gerchberg_saxton)):
    for t in targets[:4]: once(t, fn)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in targets[4:]: once(t, fn)
    torch.cuda.synchronize(); print(name, "ms per call", 1e3 * (time.perf_counter() - t0) / 20)
    pr = cProfile.Profile(); pr.enable()
    for t in targets[4:]: once(t, fn)
    pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
