"""TEST INFRASTRUCTURE ONLY -- CPU restatements of the reference algorithms.

Nothing under ``synthanatomy_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker / CPU baseline -- never as the shipped path.
"""
