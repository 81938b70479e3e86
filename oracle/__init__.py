"""CPU oracle for the UST-RUN SSL train step (TEST INFRASTRUCTURE ONLY).

This package is a torch-CPU restatement of the reference's algorithm for the hot path
(SURVEY.md section 8).  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``ust-run_b200/`` imports it; the product path fails loudly when the CUDA library is missing.

Pinning: the reference ships no tests, golden vectors or known-answer tests, so
the oracle is pinned against *outputs of the reference itself run in the build container*:
``oracle/make_golden.py`` imports ``/root/reference`` (networks/unet_model.py, networks/unet.py,
networks/dsbn.py, utils/losses.py, utils/ramps.py), runs it on seeded inputs and commits the
results under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks this restatement against
those fixtures (and live against the reference when ``/root/reference`` is present).
"""
