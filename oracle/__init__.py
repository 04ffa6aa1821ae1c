"""CPU oracle for the RobustCap fusion + kinematics hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package, and there only as the checker / the CPU baseline.
The product (``robustcap_b200``) never imports it and fails loudly when its CUDA library is missing.

It is a from-scratch restatement (torch CPU tensors, float32 by default, float64 on request) of the reference's
algorithm for this path; every function cites the reference ``file:line`` it follows.

Parity pin: the reference ships no tests, golden vectors or fixtures for this path, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by
``tests/golden/make_golden.py`` (which imports ``/root/reference`` unmodified) and committed as
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every oracle function against those files.
"""
