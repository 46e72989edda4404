"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the third-party Sinkhorn solver on Aspire's otAspire path.

PARITY UNPINNED: the arithmetic of this step lives in ``geomloss==0.2.4``
(reference ``requirements.txt:1``), which is NOT vendored in /root/reference and
cannot be installed offline.  The reference repo holds no golden vector, test or
fixture for it.  This file restates the published algorithm
of that release (``SamplesLoss("sinkhorn", p=1, debias=False, reach=None)`` ->
``sinkhorn_tensorized`` -> ``sinkhorn_loop`` -> ``sinkhorn_cost``) and is anchored
on the reference's call sites:

  * src/learning/facetid_models/pair_distances.py:68-72  (potentials=True)
  * src/learning/facetid_models/pair_distances.py:88-91  (potentials=False)
  * examples/ex_aspire_consent_multimatch.py:165-169,185-188 (release copies)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product package
(``aspire_b200``) never does.

Two entry points:
  * :class:`SamplesLoss` -- torch, same dtype as its inputs (fp32 on the reference
    path).  Shaped like the geomloss class so ``ref_shims`` can register this
    module as ``geomloss`` when importing the unmodified reference.
  * :func:`sinkhorn_np` -- numpy, any float dtype (float64 = high-precision oracle),
    explicit epsilon schedule.
"""
import numpy as np
import torch

__version__ = "0.2.4-restated"


# --------------------------------------------------------------------------- schedule
def max_diameter(x, y):
    """Bounding-box diagonal of the union of two flat point clouds [n, D] (torch)."""
    lo = torch.minimum(x.min(dim=0)[0], y.min(dim=0)[0])
    hi = torch.maximum(x.max(dim=0)[0], y.max(dim=0)[0])
    return (hi - lo).norm().item()


def epsilon_schedule(p, diameter, blur, scaling):
    """float64 schedule: [diam^p] + exp(arange(p ln diam, p ln blur, p ln scaling)) + [blur^p]."""
    steps = np.arange(p * np.log(diameter), p * np.log(blur), p * np.log(scaling))
    return [diameter ** p] + [float(np.exp(e)) for e in steps] + [blur ** p]


def fixed_length_schedule(diameter, blur, n):
    """BASELINE config 5: explicit n-entry schedule [diam] + geomspace(diam->blur, n-2, open) + [blur]."""
    mid = np.geomspace(diameter, blur, n - 2, endpoint=False)
    return [float(diameter)] + [float(e) for e in mid] + [float(blur)]


# --------------------------------------------------------------------------- torch solver
def _sq_dists(x, y):
    xx = (x * x).sum(-1).unsqueeze(2)
    yy = (y * y).sum(-1).unsqueeze(1)
    xy = torch.matmul(x, y.permute(0, 2, 1))
    return xx - 2 * xy + yy


def cost_p1(x, y):
    """geomloss p=1 cost: sqrt(clamp_min(|x|^2 - 2 x.y + |y|^2, 1e-8)), batched [B,N,M]."""
    return torch.sqrt(torch.clamp_min(_sq_dists(x, y), 1e-8))


def log_weights(w):
    lw = w.log()
    lw[w <= 0] = -100000
    return lw


def softmin(eps, C, h):
    """-eps * logsumexp_j(h_j - C_ij/eps) over the last axis; C [B,N,M], h [B,M] -> [B,N]."""
    B = C.shape[0]
    return -eps * (h.view(B, 1, -1) - C / eps).logsumexp(2).view(B, -1)


def sinkhorn_loop(a_log, b_log, C_xy, C_yx, eps_list):
    """Symmetric (Jacobi, averaged) epsilon-scaling loop + un-averaged last extrapolation.

    Returns (f on x [B,N], g on y [B,M]).
    """
    eps = eps_list[0]
    g = softmin(eps, C_yx, a_log)
    f = softmin(eps, C_xy, b_log)
    for eps in eps_list:
        g_t = softmin(eps, C_yx, a_log + f / eps)
        f_t = softmin(eps, C_xy, b_log + g / eps)
        g, f = 0.5 * (g + g_t), 0.5 * (f + f_t)
    g, f = softmin(eps, C_yx, a_log + f / eps), softmin(eps, C_xy, b_log + g / eps)
    return f, g


class SamplesLoss:
    """Restated subset of geomloss.SamplesLoss: loss='sinkhorn', p=1, reach=None, debias=False."""

    last_call = None  # (diameter, n_eps) of the most recent call, for tests

    def __init__(self, loss="sinkhorn", p=2, blur=0.05, reach=None, diameter=None, scaling=0.5,
                 debias=True, potentials=False, eps_list=None, **_unused):
        if loss != "sinkhorn" or reach is not None or debias or p != 1:
            raise NotImplementedError("oracle restates only the configuration Aspire uses")
        self.p, self.blur, self.diameter, self.scaling = p, blur, diameter, scaling
        self.potentials = potentials
        self.eps_list = eps_list  # oracle-only extension: explicit schedule (config 5)

    def __call__(self, a, x, b, y):
        B, N, D = x.shape
        C_xy, C_yx = cost_p1(x, y), cost_p1(y, x)
        if self.eps_list is not None:
            eps_list, diameter = list(self.eps_list), None
        else:
            diameter = self.diameter
            if diameter is None:
                diameter = max_diameter(x.reshape(-1, D), y.reshape(-1, D))
            eps_list = epsilon_schedule(self.p, diameter, self.blur, self.scaling)
        SamplesLoss.last_call = (diameter, len(eps_list))
        f, g = sinkhorn_loop(log_weights(a), log_weights(b), C_xy, C_yx, eps_list)
        if self.potentials:
            return f.view_as(a), g.view_as(b)
        return (a.view(B, -1) * f.view(B, -1)).sum(1) + (b.view(B, -1) * g.view(B, -1)).sum(1)


# --------------------------------------------------------------------------- numpy solver
def _lse(z, axis):
    m = z.max(axis=axis, keepdims=True)
    return np.log(np.exp(z - m).sum(axis=axis)) + np.squeeze(m, axis=axis)


def sinkhorn_np(alpha, beta, C, eps_list, dtype=np.float64):
    """Same loop in numpy on a given cost C [B,N,M]; returns (f, g, dual value)."""
    alpha, beta, C = (np.asarray(t, dtype=dtype) for t in (alpha, beta, C))
    with np.errstate(divide="ignore"):
        a = np.where(alpha > 0, np.log(np.where(alpha > 0, alpha, 1)), dtype(-100000))
        b = np.where(beta > 0, np.log(np.where(beta > 0, beta, 1)), dtype(-100000))

    def smin_x(eps, h):  # over j, h on y
        return -eps * _lse(h[:, None, :] - C / eps, axis=2)

    def smin_y(eps, h):  # over i, h on x
        return -eps * _lse(h[:, :, None] - C / eps, axis=1)

    eps = dtype(eps_list[0])
    g, f = smin_y(eps, a), smin_x(eps, b)
    for eps in eps_list:
        eps = dtype(eps)
        g_t, f_t = smin_y(eps, a + f / eps), smin_x(eps, b + g / eps)
        g, f = (g + g_t) / 2, (f + f_t) / 2
    g, f = smin_y(eps, a + f / eps), smin_x(eps, b + g / eps)
    return f, g, (alpha * f).sum(1) + (beta * g).sum(1)
