"""
Bulk effective sample size, the quantity behind the "NUTS ESS/sec" metric.

The reference reports ``ess_bulk`` through ArviZ (``az.ess(method="bulk")``,
``src/liesel/goose/summary_m.py:756-757``). ArviZ 1.1.0 is third-party and not under
/root/reference; this is a restatement of its published algorithm (Vehtari et al. 2021):
rank-normalise the pooled draws, split every chain in half, estimate autocovariances per half-chain
by FFT, combine them with the between/within variance estimator and truncate the autocorrelation
sum with Geyer's initial monotone sequence. Host-side NumPy post-processing (not on the hot path).
"""

from __future__ import annotations

import numpy as np
from scipy.special import ndtri
from scipy.stats import rankdata


def _autocov(x: np.ndarray) -> np.ndarray:
    """Autocovariance along the last axis (biased, 1/n) via FFT."""
    n = x.shape[-1]
    m = 1 << int(np.ceil(np.log2(2 * n)))
    xc = x - x.mean(axis=-1, keepdims=True)
    f = np.fft.rfft(xc, n=m, axis=-1)
    ac = np.fft.irfft(f * np.conj(f), n=m, axis=-1)[..., :n]
    return ac / n


def _ess_raw(x: np.ndarray) -> float:
    """ESS of one quantity; x is [chains, draws]."""
    n_chain, n_draw = x.shape
    if n_draw < 4:
        return float("nan")
    acov = _autocov(x)
    chain_mean = x.mean(axis=1)
    mean_var = acov[:, 0].mean() * n_draw / (n_draw - 1.0)
    var_plus = mean_var * (n_draw - 1.0) / n_draw
    if n_chain > 1:
        var_plus += chain_mean.var(ddof=1)
    if not np.isfinite(var_plus) or var_plus <= 0:
        return float("nan")
    rho = np.zeros(n_draw)
    rho_even = 1.0
    rho[0] = rho_even
    rho_odd = 1.0 - (mean_var - acov[:, 1].mean()) / var_plus
    rho[1] = rho_odd
    t = 1
    while t < n_draw - 3 and (rho_even + rho_odd) > 0.0:
        rho_even = 1.0 - (mean_var - acov[:, t + 1].mean()) / var_plus
        rho_odd = 1.0 - (mean_var - acov[:, t + 2].mean()) / var_plus
        if rho_even + rho_odd >= 0:
            rho[t + 1] = rho_even
            rho[t + 2] = rho_odd
        t += 2
    max_t = t - 2
    if rho_even > 0:
        rho[max_t + 1] = rho_even
    # Geyer's initial monotone sequence
    t = 1
    while t <= max_t - 2:
        if rho[t + 1] + rho[t + 2] > rho[t - 1] + rho[t]:
            rho[t + 1] = (rho[t - 1] + rho[t]) / 2.0
            rho[t + 2] = rho[t + 1]
        t += 2
    ess = n_chain * n_draw
    tau = -1.0 + 2.0 * rho[: max_t + 1].sum() + (rho[max_t + 1] if max_t + 1 < n_draw else 0.0)
    tau = max(tau, 1.0 / np.log10(ess))
    return float(ess / tau)


def _split(x: np.ndarray) -> np.ndarray:
    n = x.shape[1]
    h = n // 2
    return np.concatenate([x[:, :h], x[:, n - h:]], axis=0)


def ess_bulk(draws: np.ndarray) -> np.ndarray:
    """
    Bulk-ESS per parameter. ``draws`` is [num_draws, num_chains, P] (the layout
    ``BatchedNUTS.sample`` returns) or [num_draws, num_chains].
    """
    d = np.asarray(draws, dtype=np.float64)
    if d.ndim == 2:
        d = d[:, :, None]
    n_draw, n_chain, P = d.shape
    out = np.empty(P)
    for j in range(P):
        x = d[:, :, j].T  # [chains, draws]
        r = rankdata(x, method="average").reshape(x.shape)
        z = ndtri((r - 0.375) / (x.size + 0.25))
        out[j] = _ess_raw(_split(z))
    return out
