"""TEST INFRASTRUCTURE ONLY — CPU oracle for the cross_fusion hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker / CPU baseline.  The product
package ``transfusion_b200`` never imports this package and fails loudly when its CUDA
library is missing.

Parity status: the reference repository holds no tests, golden vectors or fixtures for
this path.  The oracle is therefore pinned against outputs of the
*unmodified reference module itself*, imported in the build container by
``oracle/ref_loader.py`` and frozen into ``tests/golden/*.npz`` by
``oracle/make_golden.py``.
"""
