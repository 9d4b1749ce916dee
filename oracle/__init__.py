"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's GDN hot path (SchlomoFeng/GDN), used as the
checker for the CUDA kernels.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import anything
from this package.  The product (``gdn_b200``) never imports it and has no CPU
fallback.

Pinning status: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the reference's own Python files run in the
build container (``oracle/make_golden.py`` imports ``/root/reference`` through the
PyG-1.5.0 stand-in in ``oracle/pyg_shim.py`` and writes ``tests/golden/*.npz``).
The PyG dependency itself (torch-geometric==1.5.0, reference ``install.sh:1-5``) is
absent offline and restated from its documented behaviour, so parity at that
third-party boundary is "restated, not pinned".

Modules: ``gdn_oracle`` (forward/backward, op by op), ``closed_form`` (vectorised twin of the kernel algebra),
``scoring_oracle`` (evaluate.py:48-68), ``data_oracle`` (datasets/TimeDataset.py windows, test.py loop),
``metrics_oracle`` (util/data.py:28-51 sweep, evaluate.py:101-158 summary); generators ``make_golden.py``,
``make_golden_data.py``, ``make_golden_test.py``, ``make_golden_metrics.py`` (build container only).
"""
