"""CPU oracle for the pino-locoman SQP inner loop.  TEST INFRASTRUCTURE ONLY.

This package is a plain numpy (FP64 / complex128) restatement of the algorithm the
reference runs through pinocchio.casadi + casadi.Opti + OSQP.  It exists to check
the CUDA path; the product (``pino_locoman_b200``) never imports it.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import from here.

PARITY UNPINNED: none of casadi / pinocchio / osqp is installable in the build
container and the reference ships no tests or golden vectors for this path
(SURVEY.md section 8c).  The only reference code that runs here is
``utils/gait_sequence.py``; ``tests/golden/gait_*.json`` pins the oracle's gait
schedule against it bit for bit.  Everything else is pinned by algebraic
identities (EOM identity of run_ocp.py:106-161, ABA o RNEA = id, momentum
identities), complex-step derivatives and published constants (total masses,
problem sizes).

Third-party semantics restated here (no source in /root/reference):
  * pinocchio (conda-forge, unpinned, >=2.7/3.x): URDF model building, SE(3)
    integrate/difference, RNEA/ABA with external forces, centroidal map,
    dccrba, frame velocities  -> model.py, spatial.py, rbd.py
  * casadi (unpinned, >=3.6): Opti variable/parameter/constraint ordering and
    canonical forms, exact derivatives -> ocp.py (derivatives by complex step)
  * osqp (conda-forge, unpinned, 0.6.x semantics): ADMM with Ruiz scaling
    -> osqp_admm.py
"""
