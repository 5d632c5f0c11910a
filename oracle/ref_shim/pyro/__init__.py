"""TEST-ONLY stand-in for `pyro-ppl` (pinned 1.8.6 in the reference's poetry.lock, not installed here).

Lets `/root/reference/src/usflows` be imported in the build container so the oracle restatement can be
validated against the real reference and golden vectors can be generated (oracle/make_golden.py).
It is never imported by the product package `usflows_b200` and never travels into a timed path.
Adds no arithmetic except `pyro.nn.DenseNN` (plain nn.Linear/ReLU stack, restated from pyro 1.8.6 semantics).
"""
from . import distributions, nn, infer  # noqa: F401
