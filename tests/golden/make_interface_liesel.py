"""Regenerates tests/golden/interface_liesel.json: the inputs of the reference's
tests/goose/test_interface_liesel.py:12-21 (x = jax.random.normal(PRNGKey(1337), (500,))) through the
restated JAX PRNG (oracle/jax_prng.py), together with the two log-prob values that test pins (:43-44).

  python tests/golden/make_interface_liesel.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import jax_prng  # noqa: E402

x = jax_prng.normal(jax_prng.prng_key(1337), 500)
out = {
    "source": "reference tests/goose/test_interface_liesel.py:12-44",
    "x_float32_hex": [v.tobytes().hex() for v in x],
    "pins": {"log_prob_mu_0": -719.46875, "log_prob_mu_10": -25616.779296875},
}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "interface_liesel.json")
json.dump(out, open(path, "w"))
print("wrote", path)
