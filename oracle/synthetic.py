"""Kept for the tests that import `oracle.synthetic`: the synthetic MELD-shaped batch generator is product-side
input tooling (ergm_b200/synthetic.py), not part of the oracle."""
from ergm_b200.synthetic import *  # noqa: F401,F403
from ergm_b200.synthetic import make_batch, gv1_inputs  # noqa: F401
