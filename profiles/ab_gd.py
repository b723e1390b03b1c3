"""Developer tool: GD iterations/s at 32 x 1024^2 (device resident; AB_SIZE / AB_HEIGHT / AB_BATCH / AB_LOOPS change the
workload) under a few engine switches given as NAME=VALUE pairs separated by commas on the command line, e.g.
python profiles/ab_gd.py "" SLM_PIPE_CTAS=128 SLM_GD_FORM=two_pass
Each variant runs in a child process (the switches are read once per process)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, time, numpy as np, torch
sys.path.insert(0, %r)
from spatial_light_modulator_module_b200 import host_logic as hl, synthetic
from spatial_light_modulator_module_b200.engine import Engine
alg = sys.argv[1]
import os
n = int(os.environ.get("AB_SIZE", "1024"))
shape, batch, loops = (int(os.environ.get("AB_HEIGHT", n)), n), int(os.environ.get("AB_BATCH", "32")), int(os.environ.get("AB_LOOPS", "100"))
eng = Engine(shape, "fp32", batch)
dev = torch.device("cuda", 0)
t = torch.from_numpy(np.stack([synthetic.noise_target(shape, seed=i) for i in range(batch)])).to(dev)
x0 = torch.from_numpy(np.exp(2j * np.pi * np.random.default_rng(0).random((batch,) + shape)).astype(np.complex64)).to(dev)
x = torch.empty_like(x0)
during, _ = hl.learning_rate_schedule(0.005, 0, loops)
norms = np.full(batch, 255.0)
print("", end="")
def step():
    if alg == "gd":
        x.copy_(x0); r, _ = eng.gd(t, x, during, loops, want_expected=False, norms=norms)
    else:
        r = eng.gs(t, loops, want_expected=False, norms=norms)
    return r
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): r = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
eng.profile(True); eng.profile_read(); step(); p = eng.profile_read()
print("%%8.0f it/s  %%6.2f ms/step  err %%.6g   " %% (batch * loops / ms * 1e3, ms, r.errors[0][-1]) +
      "  ".join("%%s %%.4f" %% (k, v[0] / v[1]) for k, v in p.items() if v[1] and k in ("row_pass", "col_pass", "col_stats")))
''' % ROOT
alg = "gd"
for spec in sys.argv[1:] or [""]:
    if spec in ("gd", "gs"):
        alg = spec
        continue
    env = dict(os.environ)
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("=", 1)
        env[k] = v
    out = subprocess.run([sys.executable, "-c", CODE, alg], env=env, capture_output=True, text=True)
    print(f"{alg} [{spec or 'default'}]: {out.stdout.strip() or out.stderr.strip()[-400:]}", flush=True)
