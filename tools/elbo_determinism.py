import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import vgposp_b200.gp_functions as gpf
n, m, b = int(os.environ.get("N", 200000)), 512, 4096
rng = np.random.default_rng(1)
x = rng.uniform(-2, 2, (n, 3)); y = np.sum(np.sin(2*np.pi*x), axis=1) + 0.1*rng.standard_normal(n); z = rng.uniform(-2, 2, (m, 3))
tr = gpf.VgpTrainer(x, y, z, b)
idx = rng.integers(n, size=b)
outs = [tr.loss_and_grad(x[idx], y[idx]) for _ in range(5)]
l0, g0, gz0, t0 = outs[0]
print("first call: loss %.12f grads %s |gz| %.12e gz[0] %s" % (l0, g0, np.abs(gz0).sum(), gz0[0]))
for i, (l, g, gz, t) in enumerate(outs[1:], 1):
    bad = np.where(np.abs(gz - gz0).max(axis=1) > 0)[0]
    print("rows differing", len(bad), bad[:10])
    print("ov", os.environ.get("VGP_ELBO_OVERLAP"), "call", i, "loss diff %.3e" % abs(l - l0), "grad diff", np.abs(g - g0), "gradz max diff %.3e" % np.abs(gz - gz0).max(),
          {k: abs(t[k] - t0[k]) for k in t})
