"""Host-side scene helpers: what `env.reset` hands to the hot path.

The occupancy grid (simulator/scene/scene_generator.py:292-328,888-922) is uploaded as a list
of half-open zero-cell rectangles; the robot/wall test of simulator/env.py:227-261 (window
sum over scene.map) is then an integer AABB-overlap test, bit-identical to the grid test.
"""
import numpy as np


def rects_from_zero_cells(zero):
    """Exact decomposition of a boolean grid (True = obstacle cell, scene.map == 0) into
    disjoint rectangles [x0, x1) x [y0, y1) (x = first index, y = second index).
    Greedy: maximal runs along y per row, merged along x while identical."""
    zero = np.asarray(zero, dtype=bool)
    gx, gy = zero.shape
    open_runs = {}   # (y0, y1) -> x0
    rects = []
    for x in range(gx + 1):
        runs = set()
        if x < gx:
            row = zero[x]
            if row.any():
                d = np.diff(np.concatenate(([0], row.astype(np.int8), [0])))
                starts, ends = np.nonzero(d == 1)[0], np.nonzero(d == -1)[0]
                runs = set(zip(starts.tolist(), ends.tolist()))
        for run in list(open_runs):
            if run not in runs:
                rects.append((open_runs.pop(run), run[0], x, run[1]))
        for run in runs:
            if run not in open_runs:
                open_runs[run] = x
    rects.sort()
    return np.asarray(rects, dtype=np.int16).reshape(-1, 4)


def wall_polygon(cx, cy, x_dim, y_dim, inflation=1.0):
    """The 4 vertices the reference records for a wall (scene_generator.py:269-290): counter-clockwise from the
    (+, +) corner, inflated by `inflation` (its inflation_rate_il)."""
    hx, hy = inflation * x_dim / 2.0, inflation * y_dim / 2.0
    return [(cx + hx, cy + hy), (cx - hx, cy + hy), (cx - hx, cy - hy), (cx + hx, cy - hy)]


def pack_obstacles(polygons, cap=64):
    """RVO2 addObstacle for every polygon + processObstacles (simulator/policy/orca_obstacles.py:102-107) through
    the library's host entry point ebc_pack_obstacles -> structured array of ebc_obst_vertex records.  When the
    kd-tree's edge splitting needs more than `cap` vertices, polygons are dropped from the END of the list until it
    fits (returned as the second value: how many polygons were kept)."""
    import ctypes
    from . import abi
    lib = abi.load()
    polys = [list(p) for p in polygons]
    while True:
        out = np.zeros(max(cap, 1), dtype=abi.OBST_DTYPE)
        n = ctypes.c_int32(0)
        xy = np.asarray([c for p in polys for v in p for c in v[:2]], dtype=np.float32)
        sizes = np.asarray([len(p) for p in polys], dtype=np.int32)
        rc = lib.ebc_pack_obstacles(xy.ctypes.data_as(ctypes.c_void_p), sizes.ctypes.data_as(ctypes.c_void_p), len(polys),
                                    out.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(n))
        if rc == 0:
            return out[:n.value].copy(), len(polys)
        if not polys:
            raise abi.EbcError("ebc_pack_obstacles failed on an empty polygon list")
        polys.pop()
