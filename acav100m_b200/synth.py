"""Synthetic inputs of the shapes BASELINE.json names.

Features: Gaussian mixture (means ~ N(0, spread^2 I), unit noise) -- realistic best/second-best
margins.  Cluster-id pairs: c_a ~ truncated Zipf(s), c_v = pi(c_a) with probability `p_corr` (a
fixed random permutation pi: "corresponding" clips) else uniform.

The numpy variants use the legacy ``RandomState`` stream (frozen by numpy) so golden fixtures can
regenerate their inputs from a seed; the torch variants generate directly on a device for the
benchmark sizes.
"""
import numpy as np
import torch


def gaussian_mixture(n, d, k_true, seed, spread=3.0):
    rng = np.random.RandomState(seed)
    means = (rng.standard_normal((k_true, d)) * spread).astype(np.float32)
    comp = rng.randint(0, k_true, size=n)
    x = means[comp] + rng.standard_normal((n, d)).astype(np.float32)
    return np.ascontiguousarray(x, dtype=np.float32)


def zipf_pairs(v, k, seed, s=1.1, p_corr=0.5):
    rng = np.random.RandomState(seed)
    w = 1.0 / np.arange(1, k + 1, dtype=np.float64) ** s
    cdf = np.cumsum(w / w.sum())
    ca = np.minimum(np.searchsorted(cdf, rng.random_sample(v)), k - 1)
    pi = rng.permutation(k)
    corr = rng.random_sample(v) < p_corr
    cv = np.where(corr, pi[ca], rng.randint(0, k, size=v))
    return np.stack([ca, cv], axis=1).astype(np.int64)


def gaussian_mixture_torch(n, d, k_true, seed, device, spread=3.0, chunk=1 << 18, out=None, means_seed=None):
    """fp32 [n, d] on `device`, generated in chunks so the transient stays small.  `means_seed`: draw the component
    means from their own stream (ranks of a sharded run share the mixture and differ in the rows they hold)."""
    g = torch.Generator(device=device)
    if means_seed is not None:
        g.manual_seed(means_seed)
        means = torch.randn(k_true, d, generator=g, device=device) * spread
        g.manual_seed(seed)
    else:
        g.manual_seed(seed)
        means = torch.randn(k_true, d, generator=g, device=device) * spread
    x = out if out is not None else torch.empty(n, d, dtype=torch.float32, device=device)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        comp = torch.randint(0, k_true, (hi - lo,), generator=g, device=device)
        torch.randn(hi - lo, d, generator=g, device=device, out=x[lo:hi])
        x[lo:hi] += means[comp]
    return x


def zipf_pairs_torch(v, k, seed, device, s=1.1, p_corr=0.5):
    """int64 [v, 2] on `device` (the dtype the reference's measures take, mi.py:24)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    w = 1.0 / torch.arange(1, k + 1, dtype=torch.float64, device=device) ** s
    cdf = torch.cumsum(w / w.sum(), 0)
    u = torch.rand(v, generator=g, device=device, dtype=torch.float64)
    ca = torch.searchsorted(cdf, u).clamp_(max=k - 1)
    pi = torch.randperm(k, generator=g, device=device)
    corr = torch.rand(v, generator=g, device=device) < p_corr
    cv = torch.where(corr, pi[ca], torch.randint(0, k, (v,), generator=g, device=device))
    return torch.stack([ca, cv], dim=1).to(torch.int64)
