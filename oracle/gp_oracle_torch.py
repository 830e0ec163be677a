"""TEST INFRASTRUCTURE ONLY -- torch (CPU, float64, autograd) restatement of the VGP training step.

The same formulas as oracle/gp_oracle.py (**parity unpinned**, see there), written
with differentiable torch ops so that autograd provides the reference gradients of the "reference-faithful"
training step of variational_Gaussian_process_example.py:47-102: amplitude/length-scale/noise are softplus
images of unconstrained variables (:47-61), the variational (loc, scale) are the Titsias optimum over the FULL
training set and are functions of those variables (:68-74), the loss is the minibatch `variational_loss` with
kl_weight = B/N (:96-99), and tf.train.AdamOptimizer(0.01) updates (v_amp, v_ls, v_noise, Z) (:101-102).

Only tests/ and bench.py's CPU baseline import this module.
"""
import math

import torch

LOG_2PI = math.log(2.0 * math.pi)


KERNEL = "expquad"      # module switch read by `expquad` below: "expquad", "matern32" or "matern52" (TFP's forms)


def expquad(x1, x2, amp, ls):
    """The kernel matrix of the model (ExpQuad unless KERNEL says otherwise; the name is kept for the call sites).
    The Matern forms take r through a square root whose gradient at r = 0 is defined as 0 (the kernels are smooth
    there; TFP uses a finite-gradient sqrt for the same purpose)."""
    d = x1[:, None, :] - x2[None, :, :]
    r2 = (d * d).sum(-1)
    if KERNEL == "expquad":
        return amp ** 2 * torch.exp(r2 * (-0.5 / ls ** 2))
    pos = r2 > 0
    r = torch.where(pos, torch.sqrt(torch.where(pos, r2, torch.ones_like(r2))), torch.zeros_like(r2))
    if KERNEL == "matern32":
        u = math.sqrt(3.0) * r / ls
        return amp ** 2 * (1.0 + u) * torch.exp(-u)
    if KERNEL == "matern52":
        u = math.sqrt(5.0) * r / ls
        return amp ** 2 * (1.0 + u + u * u / 3.0) * torch.exp(-u)
    raise ValueError(KERNEL)


def constrained(v_amp, v_ls, v_noise, ls_offset=1e-5):
    sp = torch.nn.functional.softplus
    return sp(v_amp), ls_offset + sp(v_ls), sp(v_noise)


def optimal_posterior(z, x, y, amp, ls, noise, jitter=1e-6):
    m = z.shape[0]
    eye = torch.eye(m, dtype=z.dtype)
    kzz = expquad(z, z, amp, ls)
    kzx = expquad(z, x, amp, ls)
    sinv = kzz + (kzx @ kzx.T) / noise + jitter * eye
    ls_ = torch.linalg.cholesky(sinv)
    u = torch.cholesky_solve((kzx @ y)[:, None], ls_)[:, 0]
    loc = kzz @ u / noise
    scale = torch.linalg.solve_triangular(ls_, kzz, upper=False).T      # S = scale scale^T = Kzz Sigma Kzz
    return loc, scale


def vgp_loss(z, loc, scale, xb, yb, amp, ls, noise, kl_weight, jitter=1e-6):
    m, b = z.shape[0], xb.shape[0]
    eye = torch.eye(m, dtype=z.dtype)
    l = torch.linalg.cholesky(expquad(z, z, amp, ls) + jitter * eye)
    kzb = expquad(z, xb, amp, ls)
    alpha = torch.cholesky_solve(loc[:, None], l)[:, 0]
    r = yb - kzb.T @ alpha
    ll = -0.5 * (r @ r) / noise - 0.5 * b * (LOG_2PI + torch.log(noise))
    c = torch.linalg.solve_triangular(l, kzb, upper=False)
    d = torch.linalg.solve_triangular(l.T, c, upper=True)
    tr1 = b * amp ** 2 - (c * c).sum()
    e = scale.T @ d
    tr2 = (e * e).sum()
    li_a = torch.linalg.solve_triangular(l, scale, upper=False)
    li_mu = torch.linalg.solve_triangular(l, loc[:, None], upper=False)[:, 0]
    logdet_k = 2.0 * torch.log(torch.diagonal(l)).sum()
    logdet_s = torch.linalg.slogdet(scale @ scale.T)[1]
    kl = 0.5 * ((li_a * li_a).sum() + li_mu @ li_mu - m + logdet_k - logdet_s)
    return -(ll - 0.5 * (tr1 + tr2) / noise - kl_weight * kl)


def training_loss(params, x, y, xb, yb, jitter=1e-6, ls_offset=1e-5):
    """params = (v_amp, v_ls, v_noise, Z); loss of one reference-faithful step."""
    v_amp, v_ls, v_noise, z = params
    amp, ls, noise = constrained(v_amp, v_ls, v_noise, ls_offset)
    loc, scale = optimal_posterior(z, x, y, amp, ls, noise, jitter)
    return vgp_loss(z, loc, scale, xb, yb, amp, ls, noise, xb.shape[0] / x.shape[0], jitter)


def loss_and_grads(v_amp, v_ls, v_noise, z, x, y, xb, yb, jitter=1e-6, ls_offset=1e-5, kernel="expquad"):
    global KERNEL
    previous, KERNEL = KERNEL, kernel
    try:
        return _loss_and_grads(v_amp, v_ls, v_noise, z, x, y, xb, yb, jitter, ls_offset)
    finally:
        KERNEL = previous


def _loss_and_grads(v_amp, v_ls, v_noise, z, x, y, xb, yb, jitter=1e-6, ls_offset=1e-5):
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)  # noqa: E731
    params = [t(v_amp).clone().requires_grad_(True), t(v_ls).clone().requires_grad_(True),
              t(v_noise).clone().requires_grad_(True), t(z).clone().requires_grad_(True)]
    loss = training_loss(params, t(x), t(y), t(xb), t(yb), jitter, ls_offset)
    grads = torch.autograd.grad(loss, params)
    return float(loss), [g.numpy().copy() for g in grads]


class TfAdamTorch:
    """tf.train.AdamOptimizer semantics (epsilon outside the bias-corrected sqrt), cf. gp_oracle.TfAdam."""

    def __init__(self, shapes, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        import numpy as np
        self.m = [np.zeros(s) for s in shapes]
        self.v = [np.zeros(s) for s in shapes]
        self.t, self.lr, self.b1, self.b2, self.eps = 0, lr, beta1, beta2, eps

    def step(self, params, grads):
        import numpy as np
        self.t += 1
        lr_t = self.lr * np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        out = []
        for i, (p, g) in enumerate(zip(params, grads)):
            self.m[i] = self.b1 * self.m[i] + (1 - self.b1) * g
            self.v[i] = self.b2 * self.v[i] + (1 - self.b2) * g * g
            out.append(p - lr_t * self.m[i] / (np.sqrt(self.v[i]) + self.eps))
        return out
