"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the OCRL slot-attention hot path.

Nothing under ``ocrl_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` use it, and only as the checker / CPU baseline -- never as the product.

Parity pinning: the reference ships no tests or golden vectors (SURVEY.md finding 0.6), so
the oracle is pinned against outputs of the reference itself, imported in the build container
by ``oracle/reference_bridge.py`` and frozen under ``tests/golden/`` by
``oracle/make_golden.py``.
"""
