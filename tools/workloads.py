"""Synthetic inputs of the BASELINE.json configs, shared by bench.py and the parity tests so
that both sides time and check the same matrices.  Host-side data generation only: no arithmetic of the path."""
import numpy as np

SEED0 = 20261018                    # seeds 20261018 + i


def length_scale_for(n):
    """l proportional to n^(-1/3): the expected neighbour count within 2 l, hence cond(Sigma), stays n-independent."""
    return 0.5 * (1000.0 / n) ** (1.0 / 3.0)


def cloud(n, seed=SEED0 + 3):
    """cfg4 / cfg5 recipe: uniform cloud in [-2, 2]^3, a = 1, nugget 1e-2.  Returns (x, amplitude, l, nugget)."""
    x = np.random.default_rng(seed).uniform(-2.0, 2.0, (n, 3))
    return x, 1.0, length_scale_for(n), 1e-2


def mesh_cloud(n=10768, seed=SEED0 + 1):
    """cfg2: VTK-mesh-shaped cloud: 24 x 24 x 19 lattice in [-2, 2]^3 (10 944 points), the last 176 dropped (the
    reference's mesh has 10 768 points, generate_main_datasets.py:5-6), every coordinate jittered by U(-0.3 h, 0.3 h)
    of its lattice spacing (breaks the exact ties of a regular grid).  Returns (x, amplitude, l, nugget)."""
    dims = (24, 24, 19)
    axes = [np.linspace(-2.0, 2.0, d) for d in dims]
    h = np.array([4.0 / (d - 1) for d in dims])
    g = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)[:n]
    rng = np.random.default_rng(seed)
    x = g + rng.uniform(-0.3, 0.3, g.shape) * h
    return x, 1.0, length_scale_for(n), 1e-2


def sin_field(n=1000, seed=SEED0):
    """cfg1: 3D_sin_wave field on a uniform cloud: y = sin(x0) sin(x1) + 0.1 + N(0, 0.01)  (3D_sin_wave.py:96-103,
    NOISE offset :57), inducing points = first 32 rows of a second uniform draw."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-2.0, 2.0, (n, 3))
    y = np.sin(x[:, 0]) * np.sin(x[:, 1]) + 0.1 + 0.1 * rng.standard_normal(n)
    z = rng.uniform(-2.0, 2.0, (n, 3))[:32]
    return x, y, z


def elbo_problem(n=200000, m=512, seed=SEED0 + 4):
    """cfg3: y = sum_i sin(2 pi x_i) + N(0, 0.01) (gp_functions.py:78-95 SIN_DENSITY = 2; data_generation.py:53-61),
    Z = m rows of an independent uniform draw."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-2.0, 2.0, (n, 3))
    y = np.sum(np.sin(2 * np.pi * x), axis=1) + 0.1 * rng.standard_normal(n)
    z = rng.uniform(-2.0, 2.0, (m, 3))
    return x, y, z


def expquad_cov_host(x, amplitude, length_scale, nugget, block=1024):
    """K(x, x) + nugget I on the host in row blocks (bounded temporaries): the CPU arm's input builder."""
    n = x.shape[0]
    cov = np.empty((n, n))
    for i in range(0, n, block):
        xi = x[i:i + block]
        d = xi[:, None, :] - x[None, :, :]                          # direct differences: no cancellation
        cov[i:i + block] = amplitude ** 2 * np.exp(np.einsum("ijk,ijk->ij", d, d) * (-0.5 / length_scale ** 2))
    cov[np.diag_indices(n)] += nugget
    return cov
