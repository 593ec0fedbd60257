"""CPU oracle for the NTN cost+gradient hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker (or
the timed CPU baseline), never as the path that is shipped.

PARITY UNPINNED.  The reference (``/root/reference``) is Java on ND4j
0.0.3.5.5.3-SNAPSHOT + jblas 1.2.3 + umass-nlp-1.0-SNAPSHOT; none of those jars
are vendored, there is no JVM in this image, and the reference holds no tests,
golden vectors or fixtures.  This oracle is therefore a
line-by-line restatement of the reference's *source* (every function cites the
file:line it follows), pinned only by (a) central finite differences of its own
cost, (b) java.util.Random known answers from the public JDK spec, and (c) a
second, independently written float32 run.  See DESIGN.md "Oracle".
"""
