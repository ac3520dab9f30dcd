"""CPU oracle for the bundle-adjustment hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or
as the timed CPU baseline -- never as the thing shipped.  The product path
(``caliscope_b200``) must not import this package.
"""
