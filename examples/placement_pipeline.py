"""The reference's main_architecture_2 flow on the B200 path, end to end, at toy size:

    1. fit a variational GP to tracer observations over (x, y, z, temperature, pressure)      main_architecture_2.py:170-260
    2. predict the tracer on a spatial grid for a set of (pressure, temperature) samples and
       take the empirical covariance between the grid locations                               :322-494, gpf.create_cov_matrix
    3. (optionally) taper it with the decay filter                                            ..._sampledistribution.py:376-394
    4. greedy mutual-information placement of k sensors                                       alg2.placement_algorithm_2
    5. leave the CSV files the plotting scripts read                                          :754-769

    python examples/placement_pipeline.py [--cover 6] [--samples 8] [--k 5] [--steps 60] [--out /tmp/vgposp_demo]

Everything numeric runs on the GPU through libvgposp.so; the module names are the drop-ins of the reference's own.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vgposp_b200.gp_functions as gpf  # noqa: E402
import vgposp_b200.placement_algorithm2 as alg2  # noqa: E402
from vgposp_b200 import cov_producer  # noqa: E402


def tracer_field(p):
    """Synthetic tracer over (x, y, z, T, P) in [0, 2]^5."""
    return np.sin(1.5 * p[:, 0] + 0.4 * p[:, 3]) * np.cos(p[:, 1]) + 0.3 * p[:, 2] * p[:, 4]


def main(cover=6, samples=8, k=5, steps=60, n_obs=4000, m=64, batch=256, beta=0.0, out=None, seed=0, quiet=False,
         nugget=0.0):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.0, 2.0, (n_obs, 5))
    y = tracer_field(x) + 0.05 * rng.standard_normal(n_obs)
    z = rng.uniform(0.0, 2.0, (m, 5))
    # 1. VGP training (variational_Gaussian_process_example.py:46-125; Adam 0.01 on softplus parameters)
    trainer = gpf.VgpTrainer(x, y, z, batch)
    losses = []
    for _ in range(steps):
        idx = rng.integers(n_obs, size=batch)
        losses.append(trainer.step(x[idx], y[idx]))
    amp, ls, noise, z_fit = trainer.parameters()
    kernel = gpf.ExponentiatedQuadratic(amp, ls)
    loc, scale = gpf.VariationalGaussianProcess.optimal_variational_posterior(kernel, z_fit, x, y, noise, as_device=True)
    trainer.close()

    # 2. tracer predictions per grid location over the (P, T) grid -> empirical covariance (SYRK on the device)
    scale_xyz = 2.0 / max(cover - 1, 1)

    def encoder(points):                   # (i0, i1, i2, p, t) -> predicted tracer; grid index -> coordinate
        q = points.copy()
        q[:, :3] *= scale_xyz
        return gpf.VariationalGaussianProcess(kernel, q, z_fit, loc, scale, noise).mean()

    cov_vv = gpf.create_cov_matrix([0, 2], [0, 2], [0, 2], [0.0, 2.0], [0.0, 2.0], cover, samples, encoder, None)
    n = cov_vv.shape[0]
    # location index = i0 + i1 I0 + i2 I0 I1 (gp_functions.py:1041-1046)
    xyz_idxs = np.array([[i % cover, (i // cover) % cover, i // (cover * cover)] for i in range(n)], dtype=np.int32)
    # 3. optional taper.  S^2 samples give a rank-deficient covariance for n > S^2 locations: placement_algorithm_2 then
    #    runs on the pseudo-inverse path and returns what the reference's np.linalg.pinv arithmetic returns -- with
    #    rank < n / 2 that is [0, 1, ..., k - 1], every delta being 0.  `nugget` > 0 is the
    #    TF-graph variant's way out (snippets_a2.py:161-163 adds 1e-6 to the diagonal): a full-rank matrix, real scores.
    if beta > 0.0:
        cov_vv = cov_producer.cov_taper(cov_vv, xyz_idxs, beta)
    cov_in = cov_vv + (nugget * np.trace(cov_vv) / n) * np.eye(n) if nugget > 0.0 else cov_vv
    # 4. placement
    alg2.PRINTS = False
    selection = alg2.placement_algorithm_2(cov_in, k)
    # 5. hand-off files
    if out:
        os.makedirs(out, exist_ok=True)
        cache = np.zeros((n, k))
        cov_producer.save_placement_csvs(out, cov_vv, xyz_idxs, cache, np.asarray(selection))
    if not quiet:
        print("ELBO loss %.2f -> %.2f; amplitude %.3f length_scale %.3f noise %.4f" % (losses[0], losses[-1], amp, ls, noise))
        print("selected locations (index: i0 i1 i2):", [(int(s), tuple(int(v) for v in xyz_idxs[s])) for s in selection])
    return {"losses": losses, "cov_vv": cov_vv, "selection": [int(s) for s in selection], "xyz_idxs": xyz_idxs}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cover", type=int, default=6)
    ap.add_argument("--samples", type=int, default=8)
    ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--beta", type=float, default=0.0)
    ap.add_argument("--nugget", type=float, default=0.0, help="relative diagonal shift (1e-6: full-rank matrix, real scores)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    main(a.cover, a.samples, a.k, a.steps, beta=a.beta, out=a.out, nugget=a.nugget)
