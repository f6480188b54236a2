"""CPU oracle for the demcoreg co-registration hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``demcoreg_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker or the
timed CPU arm.

PARITY UNPINNED: the reference (dshean/demcoreg) ships no tests, golden vectors
or fixtures, and the arithmetic of its warp / slope / aspect / robust statistics
lives in the un-vendored, un-pinned third-party packages ``pygeotools`` (HEAD of
github.com/dshean/pygeotools, no version pin: reference ``setup.py:18``) and
GDAL (C++, unpinned).  Neither is installed in this image and there is no
network, so those semantics are restated from their published behaviour in
``oracle/pygeotools_restated.py`` (one named function per rule).  Everything that
IS importable (numpy masked arrays, ``scipy.stats.binned_statistic``,
``scipy.optimize.curve_fit``, ``scipy.signal.correlate2d``) is used unmodified.

The reference's own two files are followed line by line in
``oracle/coreglib.py`` (reference ``demcoreg/coreglib.py:14-485``) and
``oracle/dem_align.py`` (reference ``demcoreg/dem_align.py:26-172, 306-357``).
"""
