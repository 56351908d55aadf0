"""CPU oracle for the StochGPMP hot path.  TEST INFRASTRUCTURE ONLY.

This package is a from-scratch CPU restatement (numpy / torch-CPU) of the
arithmetic that ``anindex/stoch_gpmp``'s ``StochGPMP.optimize()`` performs.  It is
the checker for the CUDA kernels in ``stoch_gpmp_b200/csrc``; it is *never* the
thing that is shipped or measured as the product.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Nothing under
``stoch_gpmp_b200/`` imports it (``tests/test_no_oracle_in_product.py`` enforces
that), and the product path raises when the CUDA library is missing.

Parity status:
  * everything except Panda forward kinematics is PINNED against the real
    reference: ``oracle/make_golden.py`` runs the unmodified reference from
    ``/root/reference`` in the build container and commits the input/output vectors
    under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every oracle
    function against them.
  * Panda FK: the reference delegates FK to ``torch_robotics`` (un-vendored, no
    version pin, absent offline).  ``oracle/fk.py`` restates standard URDF
    semantics over the reference's own ``panda_arm_no_gripper.urdf`` and is pinned
    only by known answers (zero pose / shipped start pose) -> "parity unpinned"
    at that boundary.

Modules
  prior.py           closed-form block-tridiagonal precision + reverse block Cholesky
  philox.py          Philox4x32-10 + Box-Muller normal stream of the CUDA sampler
  sampler.py         banded and dense ``mu + L eps``
  fk.py              serial-chain forward kinematics (Panda constants)
  costs.py           GP / start / goal-prior / occupancy-map / sphere-RBF / IS terms
  update.py          per-particle softmax + weighted-mean update
  planner.py         whole-iteration oracle over a batch of problems
  reference_port.py  dense torch-CPU restatement of the reference's own algorithm
                     (the ``cpu_baseline`` "port" timed by bench.py)
  ref_loader.py      imports the *real* reference (build container only)
  make_golden.py     regenerates tests/golden/*.npz from the real reference
"""
