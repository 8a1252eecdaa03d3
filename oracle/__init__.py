"""CPU oracle for the HierTCN hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and there only as the checker or as the
timed CPU baseline -- never on the path that produces the framework's results.

Parity status: **parity unpinned for TensorFlow op semantics** -- the reference
is a TensorFlow-1.6 graph, TensorFlow is not installable here and the
reference ships no golden vectors.  What *is* pinned: the control flow and
tensor contract of the reference's own ``model_hier.py`` / ``model_tcn.py`` /
``customized_tcn_cell.py`` / ``loss.py``, by executing those files unmodified
under ``oracle/tf_shim`` (a numpy stand-in for the handful of TF ops they
call) -- see ``oracle/make_golden.py`` and ``tests/golden/``.
"""
