"""Stand-in for the third-party `meepmeep` package (absent from /root/reference and this image).

Used ONLY by tests/golden/make_golden.py so that the reference's own hot-path files
(pytransit/models/roadrunner/*.py) can be imported and executed unmodified when generating the
golden fixtures.  See backends/numba/point2d.py.
"""
