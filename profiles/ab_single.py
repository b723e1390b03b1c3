"""ONE hologram (batch 1) under engine switches, one child process per variant: free-running ms per hologram for
GD / GS at 1024^2 x 100 iterations and GS at 512^2 x 20 (BASELINE configs 2 and 1).  A/B on ONE box."""
import os, subprocess, sys
CHILD = r'''
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from spatial_light_modulator_module_b200 import synthetic, host_logic as hl
from spatial_light_modulator_module_b200.engine import Engine
import os
B = int(os.environ.get("AB_BATCH", "1"))
out = []
for n, loops, algs in ((1024, 100, ("gd", "gs")), (512, 20, ("gs",))):
    shape = (n, n)
    eng = Engine(shape, "fp32", B)
    dev = torch.device("cuda", 0)
    t = torch.from_numpy(np.stack([synthetic.noise_target(shape, seed=1 + i) for i in range(B)])).to(dev)
    norms = np.array([float(t[i].max()) for i in range(B)])
    x0 = torch.from_numpy(np.exp(2j * np.pi * np.random.default_rng(0).random((B,) + shape)).astype(np.complex64)).to(dev)
    x = torch.empty_like(x0)
    during, _ = hl.learning_rate_schedule(0.005, 0, loops)
    def run(alg):
        if alg == "gd":
            x.copy_(x0); eng.gd(t, x, during, loops, want_expected=False, norms=norms)
        else:
            eng.gs(t, loops, want_expected=False, norms=norms)
    for alg in algs:
        for _ in range(5): run(alg)
        best = 1e9
        for rep in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(10): run(alg)
            torch.cuda.synchronize(); best = min(best, (time.perf_counter() - t0) / 10)
        out.append(f"{alg}{n} {1e3*best:.3f}")
    eng.close()
print("  ".join(out))
'''
variants = [a.split(",") if a else [] for a in (sys.argv[1:] or ["", "SLM_PDL=0", "SLM_NO_GROUPS=1", "SLM_PDL=0,SLM_NO_GROUPS=1", "SLM_NO_GRAPH=1", "SLM_NO_GRAPH=1,SLM_PDL=0"])]
for v in variants:
    env = dict(os.environ)
    env.update(dict(kv.split("=") for kv in v))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(f"{' '.join(v) or 'default':40s} {r.stdout.strip() or r.stderr[-400:]}", flush=True)
