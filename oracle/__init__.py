"""CPU oracle for the VarNet + spatial-alignment hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``spatialalignmentnetwork_b200/`` may
import this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker (or as the timed CPU baseline), never as the product.

It is a *functional restatement* (plain functions over a flat ``state_dict``
of tensors) of the reference's algorithm.  The reference's arithmetic lives in
PyTorch library calls ("where the arithmetic really lives"), so
the restatement is written against the same CPU library (``torch`` on CPU,
fp32 or fp64 selectable) but shares no code with the reference's
``nn.Module`` classes.  Each function cites the reference ``file:line`` it
follows.

Parity pin: the reference ships no tests / golden vectors, so
the oracle is pinned against outputs of the reference itself, generated in the
build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference``) and committed as ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks every oracle function against them.
"""
from . import signal, varnet, align, losses, step, gan, augment  # noqa: F401
