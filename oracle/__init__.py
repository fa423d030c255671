"""CPU oracle for the super_sac off-policy update path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (torch-CPU fp32 / numpy, explicit backward passes, no autograd
on the restated part) of the reference's update step.  Every function cites the reference
``file:line`` it follows (paths relative to the reference root, jakegrigsby/super_sac).

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs --
as the checker or the timed CPU baseline, never on the product path.  ``super_sac_b200`` never
imports it; the product raises if its CUDA library is missing.

Pinning: the reference ships no tests / golden vectors.  The oracle is pinned against
the UNMODIFIED reference executed in the build container with injected randomness
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``; live cross-check in
``tests/test_oracle_vs_reference.py`` when ``/root/reference`` is present).
"""
