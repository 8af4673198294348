"""CPU oracle for the Sobol'/Saltelli hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain NumPy restatement of the reference algorithm (SALib's
``sample/sobol.py``, ``sample/saltelli.py``, ``sample/sobol_sequence.py``,
``analyze/sobol.py``, ``test_functions/{Ishigami,Sobol_G}.py``).  Each function cites
the reference file:line it follows.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- there only as the checker or the timed CPU
baseline, never as the product.  ``salib_b200`` never imports it; the product path
raises when the CUDA library is missing.

Parity pin: the oracle is checked in ``tests/test_oracle.py`` against
  * the reference's own known-answer vectors (first 10 Joe-Kuo points,
    ``tests/test_sobol.py:24-35``; ``tests/data/model_input.txt``; the CLI golden table
    ``tests/test_cli_analyze.py:269``; ``Sobol_G.evaluate(zeros(1,8)) = 4.0583``), and
  * outputs of the unmodified reference run in the build container, frozen under
    ``tests/golden/`` by ``oracle/gen_golden.py`` (committed).
"""
