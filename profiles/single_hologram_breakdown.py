"""Per-kernel device times of ONE hologram (batch 1, 1024^2, 100 iterations): where the latency goes."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from spatial_light_modulator_module_b200 import synthetic, host_logic as hl
from spatial_light_modulator_module_b200.engine import Engine
shape = (1024, 1024)
eng = Engine(shape, "fp32", 1)
dev = torch.device("cuda", 0)
t = torch.from_numpy(synthetic.noise_target(shape, seed=1)[None]).to(dev)
norms = np.array([float(t.max())])
x0 = torch.from_numpy(np.exp(2j * np.pi * np.random.default_rng(0).random((1,) + shape)).astype(np.complex64)).to(dev)
x = torch.empty_like(x0)
during, _ = hl.learning_rate_schedule(0.005, 0, 100)
def run(alg):
    if alg == "gd":
        x.copy_(x0); eng.gd(t, x, during, 100, want_expected=False, norms=norms)
    else:
        eng.gs(t, 100, want_expected=False, norms=norms)
for alg in ("gd", "gs"):
    for _ in range(3): run(alg)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): run(alg)
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 5
    eng.profile(True); eng.profile_read()
    for _ in range(5): run(alg)
    prof = eng.profile_read(); eng.profile(False)
    print(alg, f"free-running {1e3*wall:.3f} ms per hologram;", "per launch (events around every launch):",
          {k: (round(1e3 * v[0] / v[1], 2), v[1] // 5) for k, v in prof.items() if v[1]}, "us, launches per hologram")
