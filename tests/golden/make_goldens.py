"""
Regenerates tests/golden/*.json.

1. reference_pins.json — every closed-form number the reference's own docs/tests print for this
   path, transcribed with its source (file:line under /root/reference). The reference cannot be
   imported here (needs Python >= 3.13, JAX, TFP, BlackJAX — none installed, no network), so
   these printed values are the only outputs of the reference that exist in this environment.
2. model_lm.json — inputs of the reference's kernel integration test
   (tests/goose/kernels/model_lm.py:17-31, NumPy PCG64 seed 1337, regenerable bit for bit) with
   the Gaussian log-density, gradient and Hessian at the test's start state (beta = 1,
   log_sigma = log 0.1), computed here with SciPy/NumPy closed forms INDEPENDENT of oracle/.
3. iwls_utils.json — inputs of tests/goose/test_iwls_utils.py:8-35 (same seed) with
   np.linalg.solve / scipy.stats.multivariate_normal.logpdf answers.

Run from the repo root:  python tests/golden/make_goldens.py
"""
import hashlib
import json
import os

import numpy as np
import scipy.linalg
import scipy.stats

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def reference_pins():
    return {
        "normal_logpdf": [
            {"source": "src/liesel/model/nodes.py:1024-1032 (doctest)", "dtype": "float32",
             "y": [-0.5, 0.0, 0.5], "loc": 0.0, "scale": 1.0,
             "log_prob": [-1.0439385, -0.9189385, -1.0439385]},
            {"source": "src/liesel/model/nodes.py:1049-1051 (doctest, summed)", "dtype": "float32",
             "y": [-0.5, 0.0, 0.5], "loc": 0.0, "scale": 1.0, "log_prob_sum": -3.0068154},
            {"source": "src/liesel/goose/interface.py:57-69 (doctest)", "dtype": "float32",
             "y": [0.0], "loc": 0.0, "scale": 1.0, "log_prob": [-0.9189385]},
            {"source": "src/liesel/goose/interface.py:57-69 (doctest, x=1)", "dtype": "float32",
             "y": [1.0], "loc": 0.0, "scale": 1.0, "log_prob": [-1.4189385]},
            {"source": ".github/README.md:62-80", "dtype": "float32",
             "y": [1.314, 0.861, -1.813, 0.587, -1.408], "loc": 0.0, "scale": 1.0,
             "log_prob_sum": -8.635652},
            {"source": ".github/README.md:82-87", "dtype": "float32",
             "y": [1.314, 0.861, -1.813, 0.587, -1.408], "loc": -0.5, "scale": 1.0,
             "log_prob_sum": -9.031153},
        ],
        "model_log_prob": [
            {"source": "src/liesel/model/logprob.py:36-60 (LogProb docstring)", "dtype": "float32",
             "desc": "x ~ N(0,1) prior, 2 coefficients, no data; position x=[1,2]",
             "x": [1.0, 2.0], "log_prob": -4.3378773, "grad": [-1.0, -2.0],
             "hessian": [[-1.0, 0.0], [0.0, -1.0]]},
            {"source": "tests/model/test_model.py:543-562", "dtype": "float32",
             "desc": "obs=0 ~ N(y, 1), y = 0.5", "y_value": 0.5, "log_prob": -1.0439385, "grad": -0.5},
        ],
    }


def model_lm():
    rng = np.random.default_rng(1337)  # model_lm.py:17
    n, p = 30, 2
    beta = np.ones(p)
    sigma = 0.1
    X = np.column_stack([np.ones(n), rng.uniform(size=[n, p - 1])])
    y = rng.normal(X @ beta, sigma, size=n)
    beta_ols, rss_ols, _, _ = scipy.linalg.lstsq(X, y)
    sigma_ols = np.sqrt(rss_ols / (n - p))
    log_sigma = np.log(sigma)
    # log_prob of model_lm.py:34-41: sum(norm.logpdf(y, X @ beta, exp(log_sigma)))
    mu = X @ beta
    logp = scipy.stats.norm.logpdf(y, mu, np.exp(log_sigma)).sum()
    r = (y - mu) / sigma**2
    z2 = ((y - mu) / sigma) ** 2
    grad_beta = X.T @ r
    grad_ls = (z2 - 1.0).sum()
    H = np.zeros((3, 3))
    H[:2, :2] = -(X.T @ X) / sigma**2
    H[:2, 2] = H[2, :2] = -2.0 * (X.T @ r)
    H[2, 2] = -2.0 * z2.sum()
    return {
        "source": "tests/goose/kernels/model_lm.py:17-41 (inputs: NumPy PCG64 seed 1337)",
        "X": X.tolist(), "y": y.tolist(), "sha256_X": sha(X), "sha256_y": sha(y),
        "beta": beta.tolist(), "log_sigma": float(log_sigma),
        "beta_ols": beta_ols.tolist(), "sigma_ols": float(sigma_ols),
        "log_prob": float(logp), "grad_beta": grad_beta.tolist(), "grad_log_sigma": float(grad_ls),
        "hessian": H.tolist(),
    }


def iwls_utils():
    rng = np.random.default_rng(1337)  # test_iwls_utils.py:8
    n = 5
    lhs = rng.uniform(size=(n, n))
    lhs = 0.5 * (lhs + lhs.T) + n * np.identity(n)
    rhs = rng.uniform(size=n)
    sol = np.linalg.solve(lhs, rhs)
    cov = rng.uniform(size=(n, n))
    cov = 0.5 * (cov + cov.T) + n * np.identity(n)
    mean = rng.uniform(size=n)
    x = rng.uniform(size=n)
    lp = scipy.stats.multivariate_normal.logpdf(x, mean, cov)
    return {
        "source": "tests/goose/test_iwls_utils.py:8-35 (inputs: NumPy PCG64 seed 1337)",
        "solve": {"lhs": lhs.tolist(), "rhs": rhs.tolist(), "x": sol.tolist()},
        "mvn_log_prob": {"cov": cov.tolist(), "mean": mean.tolist(), "x": x.tolist(),
                         "log_prob": float(lp)},
    }


if __name__ == "__main__":
    for name, fn in [("reference_pins", reference_pins), ("model_lm", model_lm),
                     ("iwls_utils", iwls_utils)]:
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(fn(), f, indent=1)
        print("wrote", name)
