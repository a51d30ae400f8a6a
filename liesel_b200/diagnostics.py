"""
Bulk effective sample size, the quantity behind the "NUTS ESS/sec" metric.

The reference reports ``ess_bulk`` through ArviZ (``az.ess(method="bulk")``,
``src/liesel/goose/summary_m.py:756-757``). ArviZ 1.1.0 is third-party and not under
/root/reference; this is a restatement of its published algorithm (Vehtari et al. 2021):
rank-normalise the pooled draws, split every chain in half, estimate autocovariances per half-chain
by FFT, combine them with the between/within variance estimator and truncate the autocorrelation
sum with Geyer's initial monotone sequence.

Two implementations of the same algorithm: NumPy on the host (``ess_bulk``, ``split_rhat``) and
torch on the device that holds the draws (``ess_bulk_device``, ``split_rhat_device``):
with 1024 chains x thousands of draws x P coefficients the pooled rank transform (a sort of
chains*draws values per coefficient) and the per-chain FFT autocovariances are the expensive part
and stay on the GPU next to the chain storage; only the [P x draws] chain-averaged autocovariance
goes to the host for the (sequential, O(draws)) Geyer truncation.
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
    cm_var = x.mean(axis=1).var(ddof=1) if n_chain > 1 else 0.0
    return _ess_from_moments(acov.mean(axis=0), cm_var, n_chain, n_draw)


def _ess_from_moments(acov_mean: np.ndarray, chain_mean_var: float, n_chain: int, n_draw: int) -> float:
    """ESS from the chain-averaged autocovariance [draws] and the variance of the chain means."""
    mean_var = acov_mean[0] * n_draw / (n_draw - 1.0)
    var_plus = mean_var * (n_draw - 1.0) / n_draw
    if n_chain > 1:
        var_plus += chain_mean_var
    if not np.isfinite(var_plus) or var_plus <= 0:
        return float("nan")
    rho = np.zeros(n_draw)
    rho_even = 1.0
    rho[0] = rho_even
    rho_odd = 1.0 - (mean_var - acov_mean[1]) / var_plus
    rho[1] = rho_odd
    t = 1
    while t < n_draw - 3 and (rho_even + rho_odd) > 0.0:
        rho_even = 1.0 - (mean_var - acov_mean[t + 1]) / var_plus
        rho_odd = 1.0 - (mean_var - acov_mean[t + 2]) / var_plus
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


def _rhat(z: np.ndarray) -> float:
    """Potential scale reduction of split chains z [chains, draws] (Gelman-Rubin, ddof=1)."""
    n = z.shape[1]
    W = z.var(axis=1, ddof=1).mean()
    B_over_n = z.mean(axis=1).var(ddof=1)
    return float(np.sqrt(((n - 1.0) / n * W + B_over_n) / W))


def _z_scale(x: np.ndarray) -> np.ndarray:
    r = rankdata(x, method="average").reshape(x.shape)
    return ndtri((r - 0.375) / (x.size + 0.25))


def split_rhat(draws: np.ndarray) -> np.ndarray:
    """
    Rank-normalised split-R-hat per parameter (ArviZ ``rhat(method="rank")``, the default the
    reference's summary reports, ``goose/summary_m.py:758-759``): the larger of the bulk value
    (z-scaled split chains) and the tail value (z-scaled absolute deviations from the median).
    """
    d = np.asarray(draws, dtype=np.float64)
    if d.ndim == 2:
        d = d[:, :, None]
    out = np.empty(d.shape[2])
    for j in range(d.shape[2]):
        s = _split(d[:, :, j].T)
        bulk = _rhat(_z_scale(s))
        tail = _rhat(_z_scale(np.abs(s - np.median(s))))
        out[j] = max(bulk, tail)
    return out


# ---------------------------------------------------------------------------------------
# device versions (torch; run where the draws live)
# ---------------------------------------------------------------------------------------
def _z_scale_t(x):
    """Average-rank normal scores of ALL entries of x (any shape), ties averaged like rankdata."""
    import torch

    flat = x.reshape(-1)
    N = flat.numel()
    vals, idx = torch.sort(flat, stable=True)
    _, inv, counts = torch.unique_consecutive(vals, return_inverse=True, return_counts=True)
    last = torch.cumsum(counts, 0).to(torch.float64)  # rank of the last member of each tie group
    avg = last - (counts.to(torch.float64) - 1.0) / 2.0
    ranks = torch.empty(N, dtype=torch.float64, device=x.device)
    ranks[idx] = avg[inv]
    return torch.special.ndtri((ranks - 0.375) / (N + 0.25)).reshape(x.shape)


def _split_t(x):
    import torch

    n = x.shape[1]
    h = n // 2
    return torch.cat([x[:, :h], x[:, n - h:]], dim=0)


def ess_bulk_device(draws) -> np.ndarray:
    """``ess_bulk`` for a torch tensor [num_draws, num_chains, P] on its own device."""
    import torch

    d = draws.to(torch.float64)
    if d.dim() == 2:
        d = d[:, :, None]
    P = d.shape[2]
    out = np.empty(P)
    for j in range(P):
        z = _split_t(_z_scale_t(d[:, :, j].T.contiguous()))  # [2*chains, draws/2]
        n_chain, n_draw = z.shape
        if n_draw < 4:
            out[j] = float("nan")
            continue
        m = 1 << int(np.ceil(np.log2(2 * n_draw)))
        zc = z - z.mean(dim=1, keepdim=True)
        f = torch.fft.rfft(zc, n=m, dim=1)
        acov_mean = (torch.fft.irfft(f * f.conj(), n=m, dim=1)[:, :n_draw] / n_draw).mean(dim=0)
        cm_var = float(z.mean(dim=1).var(unbiased=True)) if n_chain > 1 else 0.0
        out[j] = _ess_from_moments(acov_mean.cpu().numpy(), cm_var, n_chain, n_draw)
    return out


def split_rhat_device(draws) -> np.ndarray:
    """``split_rhat`` for a torch tensor [num_draws, num_chains, P] on its own device."""
    import torch

    d = draws.to(torch.float64)
    if d.dim() == 2:
        d = d[:, :, None]

    def rhat(z):
        n = z.shape[1]
        W = z.var(dim=1, unbiased=True).mean()
        B_over_n = z.mean(dim=1).var(unbiased=True)
        return float(torch.sqrt(((n - 1.0) / n * W + B_over_n) / W))

    out = np.empty(d.shape[2])
    for j in range(d.shape[2]):
        s = _split_t(d[:, :, j].T.contiguous())
        srt = torch.sort(s.reshape(-1)).values  # np.median: mean of the two middle values
        med = 0.5 * (srt[(srt.numel() - 1) // 2] + srt[srt.numel() // 2])
        out[j] = max(rhat(_z_scale_t(s)), rhat(_z_scale_t((s - med).abs())))
    return out
