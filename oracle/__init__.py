"""
CPU oracle for the B200 hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy (fp64 by default) restatement of the reference's algorithm for the path named in
``BASELINE.json: north_star``. Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it. Nothing under
``liesel_b200/`` imports it; the product path fails loudly when the CUDA library is
missing instead of falling back to this code.

Pinning status. The reference (liesel 0.5.1) is pure Python on top of JAX, TFP and
BlackJAX; none of those is installed in the build container or on the GPU image and
Python >= 3.13 is required, so the reference itself cannot be imported or run here.
The oracle is therefore pinned against
  * every closed-form number the reference's own docs and tests print for this path
    (``tests/golden/reference_pins.json``; sources cited per entry),
  * the NumPy/SciPy baselines the reference's tests compare to
    (``tests/distributions/test_mvn_degen.py:27-78``, ``tests/goose/test_iwls_utils.py:11-35``,
    ``tests/contrib/test_splines.py:48-59``), and
  * SciPy's independent implementations of the third-party densities (TFP is not vendored
    under /root/reference): ``scipy.stats.{norm,poisson,bernoulli,invgamma}``.
DistReg-model log-prob / gradient / information values are NOT pinned by any reference
test ("parity unpinned" for those); for them the oracle is cross-checked by
finite differences in fp64.
"""

from . import densities, iwls, sam, splines  # noqa: F401
