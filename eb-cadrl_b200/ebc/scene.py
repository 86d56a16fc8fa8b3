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
