"""TEST INFRASTRUCTURE ONLY (oracle). Stand-ins for the two third-party packages the reference's pure-Python
DB branch imports and this image lacks, so that the UNMODIFIED `DBPostProcess.boxes_from_bitmap`
(R/pytocr/postprocess/db_postprocess.py:76-194, cpp_speedup=False) can run in the authoring container:

  * `pyclipper` (R/requirements.txt: pyclipper==1.1.0.post3) - a Cython wrapper of Clipper 6.4.2. The shim
    forwards `PyclipperOffset().AddPath(path, JT_ROUND, ET_CLOSEDPOLYGON)` + `Execute(delta)`
    (db_postprocess.py:147-149) to the reference's OWN vendored clipper.cpp compiled by oracle/build_ref.py
    (oracle/_ref/libclipper_ref.so, same library version, default MiterLimit 2.0 / ArcTolerance 0.25 as
    pyclipper's). pyclipper converts every coordinate to Clipper's 64-bit `cInt` through Cython's object ->
    C integer coercion (`IntPoint(py_point[0], py_point[1])` in `_to_clipper_point`), which calls the
    number's `__int__`: truncation toward zero, the same as the C++ branch's `int(box[i][0])`
    (db_postprocess.cpp:43-46). `Execute` returns a list of paths, each a list of `[x, y]` Python ints.
  * `shapely.geometry.Polygon` (shapely==1.7.0 over GEOS): only `.area` and `.length` of a closed ring are
    used (db_postprocess.py:145-146). Restated from GEOS' published algorithms in float64:
    `Area::ofRingSigned` (coordinates shifted by x0, sum of x_i * (y_{i-1} - y_{i+1}), halved, absolute
    value) and `Length::ofLine` (sum of hypot over the closed ring).

Installed into `sys.modules` by `install()`; used by tests/golden/make_golden.py only.
"""
import math
import sys
import types

import numpy as np

JT_SQUARE, JT_ROUND, JT_MITER = 0, 1, 2
ET_CLOSEDPOLYGON, ET_CLOSEDLINE, ET_OPENBUTT, ET_OPENSQUARE, ET_OPENROUND = 0, 1, 2, 3, 4


class PyclipperOffset(object):
    def __init__(self, miter_limit=2.0, arc_tolerance=0.25):
        assert miter_limit == 2.0 and arc_tolerance == 0.25, "shim covers Clipper's defaults only"
        self._paths = []

    def AddPath(self, path, join_type, end_type):
        assert join_type == JT_ROUND and end_type == ET_CLOSEDPOLYGON, "shim covers the DB call only"
        self._paths.append([(int(p[0]), int(p[1])) for p in path])      # __int__: truncation toward zero

    def Execute(self, delta):
        from . import db_oracle
        assert len(self._paths) == 1
        assert db_oracle._clipper() is not None, "needs oracle/_ref/libclipper_ref.so (oracle/build_ref.py)"
        return [[[int(x), int(y)] for x, y in p] for p in db_oracle.clipper_offset(self._paths[0], float(delta))]


def ring_area(pts):
    """GEOS Area::ofRingSigned on the closed ring, absolute value (shapely Polygon.area, no holes)."""
    p = [(float(x), float(y)) for x, y in pts]
    if len(p) < 3:
        return 0.0
    if p[0] != p[-1]:
        p.append(p[0])
    x0, s = p[0][0], 0.0
    for i in range(1, len(p) - 1):
        s += (p[i][0] - x0) * (p[i - 1][1] - p[i + 1][1])
    return abs(s / 2.0)


def ring_length(pts):
    """GEOS Length::ofLine on the closed exterior ring (shapely Polygon.length, no holes)."""
    p = [(float(x), float(y)) for x, y in pts]
    if p[0] != p[-1]:
        p.append(p[0])
    s = 0.0
    for i in range(1, len(p)):
        dx, dy = p[i][0] - p[i - 1][0], p[i][1] - p[i - 1][1]
        s += math.sqrt(dx * dx + dy * dy)
    return s


class Polygon(object):
    def __init__(self, shell):
        self._pts = np.asarray(shell).reshape(-1, 2)

    @property
    def area(self):
        return ring_area(self._pts)

    @property
    def length(self):
        return ring_length(self._pts)


def install():
    """Registers the shims as `pyclipper`, `shapely`, `shapely.geometry`; restores `np.int` (removed in
    numpy 1.24; the reference pins numpy 1.19.5 and calls `.astype(np.int)`, db_postprocess.py:186-189)."""
    pc = types.ModuleType("pyclipper")
    for k in ("JT_SQUARE", "JT_ROUND", "JT_MITER", "ET_CLOSEDPOLYGON", "ET_CLOSEDLINE", "ET_OPENBUTT",
              "ET_OPENSQUARE", "ET_OPENROUND"):
        setattr(pc, k, globals()[k])
    pc.PyclipperOffset = PyclipperOffset
    sh, geo = types.ModuleType("shapely"), types.ModuleType("shapely.geometry")
    geo.Polygon = Polygon
    sh.geometry = geo
    sys.modules["pyclipper"], sys.modules["shapely"], sys.modules["shapely.geometry"] = pc, sh, geo
    if not hasattr(np, "int"):
        np.int = int
