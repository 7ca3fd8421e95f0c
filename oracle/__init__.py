"""CPU oracle for the KNN hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker (or
as the CPU arm being timed).  Nothing under ``clip_database_b200/`` imports it.

PARITY UNPINNED: the reference (``/root/reference/image_database.py``) delegates
the distance arithmetic to the third-party ``sqlite-vec`` extension
(``requirements.txt:6``, unpinned), which is not vendored and not installed in
this image, and it ships no tests or golden vectors.  The oracle restates
sqlite-vec's published scalar algorithm and is anchored on the reference's own
call sites and on the real SQLite in this image (see ``sql_harness.py``).
"""
