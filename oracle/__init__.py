"""CPU oracle for the point-set-distance hot path. TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package, and only as the checker; ``fpsg_b200`` never does. See ``fpsg_oracle.c`` for parity status
(Chamfer: parity unpinned -- Kaolin 0.9.0 is absent; kNN: pinned against the reference's own function).
"""
