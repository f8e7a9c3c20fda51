"""CPU oracle for the UpliftingTableTennis inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the CPU arm being timed), never as the thing shipped.

Every function restates, in plain numpy / CPU torch, the arithmetic of one
reference function and cites the reference ``file:line`` it follows.  The
restatement is *pinned*: ``oracle/gen_golden.py`` imports the real reference
from ``/root/reference`` (in the build container, where it exists), runs both on
identical seeded inputs and writes the reference's outputs to
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` then checks the oracle
against those files everywhere (including the GPU box, where the reference
itself is absent).
"""
