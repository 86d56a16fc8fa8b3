"""Synthetic scenes of the BASELINE.json shapes, for throughput runs and tests.

Counter-based: every random number is splitmix64(seed, global episode id, stream, draw), so
episode e's scene does not depend on how episodes are sharded over ranks or on the batch size.
The distribution follows simulator/scene/scene_generator.py (square_crossing :672-712,
circle_crossing :593-618, walls :205-290, static discs :380-422, grid :888-922); it does NOT
replay numpy's MT19937 stream (that is the host parity path, ebc/scene.py).
"""
import numpy as np

SEED_BASE = 20261018
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(x):
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return x ^ (x >> np.uint64(31))


def uniform(seed, episode, stream, draw):
    """U[0,1) double from (seed, episode id, stream id, draw index); all broadcastable integers."""
    with np.errstate(over="ignore"):
        h = _mix(np.uint64(seed) ^ _mix(np.asarray(episode, dtype=np.uint64)))
        h = _mix(h ^ (np.asarray(stream, dtype=np.uint64) * np.uint64(0xD6E8FEB86659FD93)))
        h = _mix(h ^ (np.asarray(draw, dtype=np.uint64) * np.uint64(0xCA5A826395121157)))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


class SceneShape(object):
    """Shape of a synthetic workload.  types: list of (type code, count, v_pref range, radius range)."""

    def __init__(self, name, types, rule="square_crossing", square_width=11.0, circle_radius=5.0,
                 robot_radius=0.3, robot_v_pref=0.6, num_walls=3, wall_len=(3, 3), map_size_m=9.0,
                 map_resolution=0.1, discomfort_dist=0.1):
        self.name, self.types, self.rule = name, types, rule
        self.square_width, self.circle_radius = square_width, circle_radius
        self.robot_radius, self.robot_v_pref = robot_radius, robot_v_pref
        self.num_walls, self.wall_len = num_walls, wall_len
        self.map_size_m, self.map_resolution = map_size_m, map_resolution
        self.discomfort_dist = discomfort_dist

    @property
    def H(self):
        return sum(t[1] for t in self.types)

    @property
    def discs_per_wall(self):
        # scene_generator.py:380-422: discs of radius half_thickness*sqrt(2) stepping 2r while < far edge
        rr, pos, k = 0.5 * np.sqrt(2.0), 0.5 * np.sqrt(2.0), 0
        while pos < self.wall_len[1]:
            k, pos = k + 1, pos + 2.0 * rr
        return k

    @property
    def Smax(self):
        return self.num_walls * self.discs_per_wall

    @property
    def Rmax(self):
        return max(self.num_walls, 1)

    def to_abi(self, max_tries=64):
        """ebc_scene_shape for the device generator (ebc_generate): same distribution, same draws as generate()."""
        from . import abi
        s = abi.EbcSceneShape()
        s.n_types = len(self.types)
        for i, (code, count, (v0, v1), (r0, r1)) in enumerate(self.types):
            s.type_code[i], s.type_count[i] = code, count
            s.v_pref_lo[i], s.v_pref_hi[i], s.radius_lo[i], s.radius_hi[i] = v0, v1, r0, r1
        s.rule = {"square_crossing": 0, "circle_crossing": 1, "mixed": 2, "mixed_20": 3, "one_static": 4}[self.rule]
        s.num_walls, s.wall_len_lo, s.wall_len_hi = self.num_walls, self.wall_len[0], self.wall_len[1]
        s.discs_per_wall, s.max_tries = self.discs_per_wall, max_tries
        s.square_width, s.circle_radius = self.square_width, self.circle_radius
        s.robot_radius, s.robot_v_pref = self.robot_radius, self.robot_v_pref
        s.map_size_m, s.map_resolution, s.discomfort_dist = self.map_size_m, self.map_resolution, self.discomfort_dist
        return s


# BASELINE.json configs[1]: ranges from data/eb-cadrl/adults_8_..._fix_static.config:61-88
CFG2 = SceneShape("cfg2_h10_typed_3walls", [(0, 4, (0.4, 0.8), (0.2, 0.4)), (1, 3, (0.5, 1.5), (0.3, 0.6)),
                                           (2, 3, (0.3, 1.0), (0.1, 0.4))])
# configs[2]: unicycle robot, 5 adults on a circle, no walls
CFG3 = SceneShape("cfg3_h5_circle_unicycle", [(0, 5, (1.0, 1.0), (0.3, 0.3))], rule="circle_crossing",
                  circle_radius=4.0, robot_v_pref=1.0, num_walls=0)
# configs[3]: dense crowd, 20 humans + 10 walls on a 14 m map
CFG4 = SceneShape("cfg4_h20_10walls", [(0, 20, (1.0, 1.0), (0.3, 0.3))], rule="mixed", square_width=13.0,
                  circle_radius=6.0, robot_v_pref=1.0, num_walls=10, wall_len=(2, 4), map_size_m=14.0)
# the reference's full `mixed_20` rule (scene_generator.py:577-582): randint(20) static adults + the dynamic mix
MIXED20 = SceneShape("mixed_20_h20_5walls", [(0, 20, (0.5, 1.5), (0.2, 0.4))], rule="mixed_20", square_width=13.0,
                     circle_radius=6.0, robot_v_pref=1.0, num_walls=5, wall_len=(2, 4), map_size_m=14.0)
# configs/env_configs/env_one_static_human.config: the `one_static` rule (scene_generator.py:583-589), no obstacles
ONE_STATIC = SceneShape("one_static_h2", [(0, 2, (1.0, 1.0), (0.3, 0.3))], rule="one_static", square_width=10.0,
                        circle_radius=4.0, robot_v_pref=1.0, num_walls=0, discomfort_dist=0.2)
# configs[0] shape (the reference's CPU-runnable case): 5 adults on a circle of radius 3
CFG1 = SceneShape("cfg1_h5_circle", [(0, 5, (0.6, 0.6), (0.2, 0.2))], rule="circle_crossing", square_width=9.0,
                  circle_radius=3.0, robot_radius=0.2, robot_v_pref=0.7, num_walls=0)


def wall_geometry(lx, ly, xd, yd, G, res, discs_per_wall):
    """What the reference derives from a wall (grid location lx, ly relative to the map centre; dimensions xd, yd in
    metres), vectorised over episodes:
      * its zero-cell rectangle in scene.map (scene_generator.py:255-268,888-922): dim = round(d / res) cells,
        location = round(l + G / 2); a wall strictly inside the map is the slice [start, start + dim) with
        start = round(location - dim / 2); one that touches the border goes through the per-cell path, which only
        writes cells 0 < index < G -- so the rectangle is [max(start, 1), min(start + dim, G)) on both axes either way
        (dim / 2 is a whole number of cells for whole-metre walls on a 0.1 m grid);
      * its static discs (:380-422): radius = half thickness * sqrt(2), centres from (low edge + r) in steps of 2 r
        while < high edge; a square wall is one disc at its centre.
    -> (rect [N, 4] int16, list of (x, y, r, live) per disc slot)"""
    dimx = np.rint(xd / res).astype(np.int64)
    dimy = np.rint(yd / res).astype(np.int64)
    locx = np.rint(lx + G / 2.0).astype(np.int64)
    locy = np.rint(ly + G / 2.0).astype(np.int64)
    x0 = np.rint(locx - dimx / 2.0).astype(np.int64)
    y0 = np.rint(locy - dimy / 2.0).astype(np.int64)
    r = np.stack([np.clip(x0, 1, G), np.clip(y0, 1, G), np.clip(x0 + dimx, 1, G), np.clip(y0 + dimy, 1, G)], 1).astype(np.int16)
    xm, ym = lx * res, ly * res
    square = xd == yd
    horiz = xd > yd
    half_t = np.where(horiz, yd, xd) / 2.0
    rr = half_t * np.sqrt(2.0)
    lo = np.where(horiz, xm - xd / 2.0, ym - yd / 2.0)
    hi = np.where(horiz, xm + xd / 2.0, ym + yd / 2.0)
    pos = lo + rr
    discs = []
    for k in range(discs_per_wall):
        live = np.where(square, k == 0, pos < hi)
        sx = np.where(square, xm, np.where(horiz, pos, xm))
        sy = np.where(square, ym, np.where(horiz, ym, pos))
        discs.append((sx, sy, rr, live))
        pos = pos + 2.0 * rr
    return r, discs


def generate(shape, episode_ids, seed=SEED_BASE, max_tries=64, max_obst=0):
    """-> dict of numpy arrays in the ebc_state layout for the given global episode ids.  max_obst > 0 adds the
    walls as ORCA obstacle polygons (`obst`, `obst_count`: ebc_obst_vertex records)."""
    ids = np.asarray(episode_ids, dtype=np.uint64)
    N, H = len(ids), shape.H
    Smax, Rmax = max(shape.Smax, 1), shape.Rmax
    hum_pv = np.zeros((N, H, 4), np.float32)
    hum_gr = np.zeros((N, H, 4), np.float32)
    hum_type = np.zeros((N, H), np.uint8)
    rob_pv = np.zeros((N, 4), np.float32)
    rob_gr = np.zeros((N, 4), np.float32)
    R = shape.circle_radius
    rob_pv[:, 1] = -R
    rob_gr[:, 1] = R
    rob_gr[:, 2] = shape.robot_v_pref
    rob_gr[:, 3] = shape.robot_radius
    types = np.concatenate([np.full(c, t, np.uint8) for t, c, _, _ in shape.types])
    hum_type[:] = types[None]
    px = np.zeros((N, H)); py = np.zeros((N, H)); gx = np.zeros((N, H)); gy = np.zeros((N, H))
    rad = np.zeros((N, H)); vpref = np.zeros((N, H))
    h0 = 0
    for t, c, (v0, v1), (r0, r1) in shape.types:
        for k in range(c):
            h = h0 + k
            vpref[:, h] = v0 + (v1 - v0) * uniform(seed, ids, 1000 + h, 0)
            rad[:, h] = r0 + (r1 - r0) * uniform(seed, ids, 1000 + h, 1)
        h0 += c
    hw = shape.square_width / 2.0
    # rule mixed_20 (scene_generator.py:577-582): randint(20) static adults in a 6 x 8 box, the rest dynamic
    n_static = np.floor(uniform(seed, ids, 4000, 0) * 20.0).astype(np.int64) if shape.rule == "mixed_20" else np.zeros(N, np.int64)
    n_dynamic = H - n_static
    for h in range(H):
        first_of_type = int(np.nonzero(types == types[h])[0][0])
        is_static_all = h < n_static
        hd = h - n_static
        circle_all = ~is_static_all & ((shape.rule == "circle_crossing") | ((shape.rule in ("mixed", "mixed_20")) & (hd < n_dynamic // 2)))
        fixed = is_static_all & (h == 0)                      # :463-466: the first static adult is fixed
        px[fixed, h], py[fixed, h], gx[fixed, h], gy[fixed, h] = -0.5, -2.5, -0.5, -2.5
        if shape.rule == "one_static":                        # :583-589: two adults standing at (-2, -8) and (-3, -8)
            assert H == 2 and len(shape.types) == 1
            fixed = np.ones(N, bool)
            px[:, h], py[:, h], gx[:, h], gy[:, h] = -2.0 - h, -8.0, -2.0 - h, -8.0
        todo = ~fixed
        sign_all = np.where(uniform(seed, ids, 2000 + h, 7) < 0.5, 1.0, -1.0)
        md_robot_all = rad[:, h] + shape.robot_radius + shape.discomfort_dist
        for attempt in range(max_tries):
            if not todo.any():
                break
            e = ids[todo]
            u = lambda d: uniform(seed, e, 2000 + h, attempt * 8 + d)  # noqa: E731
            st_, ci_ = is_static_all[todo], circle_all[todo]
            md_robot = md_robot_all[todo]
            u0, u1, u2 = u(0), u(1), u(2)
            # static adult (:467-487)
            xs = u0 * 6.0 * 0.5 * sign_all[todo]
            ys = (u1 - 0.5) * 8.0
            # circle crossing (:593-618)
            ang = u0 * 2.0 * np.pi
            xc, yc = R * np.cos(ang), R * np.sin(ang)
            # square crossing (:672-712): top, bottom, left, right
            side = np.floor(u0 * 4.0).astype(np.int64)
            a = -hw + 2.0 * hw * u1
            b = -hw + 2.0 * hw * u2
            xq = np.where(side == 0, a, np.where(side == 1, a, np.where(side == 2, -hw, hw)))
            yq = np.where(side == 0, hw, np.where(side == 1, -hw, a))
            txq = np.where(side == 0, b, np.where(side == 1, b, np.where(side == 2, hw, -hw)))
            tyq = np.where(side == 0, -hw, np.where(side == 1, hw, b))
            x = np.where(st_, xs, np.where(ci_, xc, xq))
            y = np.where(st_, ys, np.where(ci_, yc, yq))
            tx = np.where(st_, xs, np.where(ci_, -xc, txq))
            ty = np.where(st_, ys, np.where(ci_, -yc, tyq))
            # reject starts too close to the robot or to agents of the same list placed earlier (all rules); the circle
            # rule also keeps clear of everybody's GOAL (:608-616), static adults of the robot's goal (:480-483)
            ok = np.hypot(x - 0.0, y + R) >= md_robot
            ok &= ~ci_ | (np.hypot(x - 0.0, y - R) >= md_robot)
            for j in range(first_of_type, h):
                mdj = rad[todo, h] + rad[todo, j] + shape.discomfort_dist
                ok &= np.hypot(x - px[todo, j], y - py[todo, j]) >= mdj
                ok &= ~ci_ | (np.hypot(x - gx[todo, j], y - gy[todo, j]) >= mdj)
            last_r = rad[todo, h - 1] if h > first_of_type else np.full(len(e), shape.robot_radius)
            ok &= ~st_ | (np.hypot(x - 0.0, y - R) >= rad[todo, h] + last_r + shape.discomfort_dist)
            if attempt == max_tries - 1:
                ok[:] = True
            idx = np.nonzero(todo)[0][ok]
            px[idx, h], py[idx, h], gx[idx, h], gy[idx, h] = x[ok], y[ok], tx[ok], ty[ok]
            todo[idx] = False
    hum_pv[:, :, 0], hum_pv[:, :, 1] = px, py
    hum_gr[:, :, 0], hum_gr[:, :, 1], hum_gr[:, :, 2], hum_gr[:, :, 3] = gx, gy, vpref, rad
    # walls: scene_generator.py:205-290 -> grid rectangle (:888-922) + static discs (:380-422)
    G = int(round(shape.map_size_m / shape.map_resolution))
    res = shape.map_resolution
    stat = np.zeros((N, Smax, 4), np.float32)
    stat_count = np.zeros(N, np.int32)
    rect = np.zeros((N, Rmax, 4), np.int16)
    rect_count = np.zeros(N, np.int32)
    clear = shape.robot_radius + shape.discomfort_dist
    walls = [[] for _ in range(N)]
    for w in range(shape.num_walls):
        todo = np.ones(N, bool)
        lx = np.zeros(N); ly = np.zeros(N); xd = np.ones(N); yd = np.ones(N)
        for attempt in range(max_tries):
            if not todo.any():
                break
            e = ids[todo]
            u = lambda d: uniform(seed, e, 3000 + w, attempt * 8 + d)  # noqa: E731
            cx = np.floor(-G / 2.0 + G * u(0))
            cy = np.floor(-G / 2.0 + G * u(1))
            length = np.floor(shape.wall_len[0] + (shape.wall_len[1] - shape.wall_len[0] + 1) * u(3))
            horiz = u(2) > 0.5
            xdim = np.where(horiz, length, 1.0)
            ydim = np.where(horiz, 1.0, length)
            xm, ym = cx * res, cy * res
            near_start = (np.abs(xm - 0.0) < xdim / 2 + clear) & (np.abs(ym + R) < ydim / 2 + clear)
            near_goal = (np.abs(xm - 0.0) < xdim / 2 + clear) & (np.abs(ym - R) < ydim / 2 + clear)
            ok = ~(near_start | near_goal)
            if attempt == max_tries - 1:
                ok[:] = True
            idx = np.nonzero(todo)[0][ok]
            lx[idx], ly[idx], xd[idx], yd[idx] = cx[ok], cy[ok], xdim[ok], ydim[ok]
            todo[idx] = False
        r, discs = wall_geometry(lx, ly, xd, yd, G, res, shape.discs_per_wall)
        rect[:, w] = r
        rect_count += 1
        xm, ym = lx * res, ly * res
        if max_obst:
            for e in range(N):
                walls[e].append((xm[e], ym[e], xd[e], yd[e]))
        for sx, sy, rr, live in discs:
            k = stat_count[live]
            rows = np.nonzero(live)[0]
            stat[rows, k, 0] = sx[live]
            stat[rows, k, 1] = sy[live]
            stat[rows, k, 2] = rr[live]
            stat_count[live] += 1
    extra = {}
    if max_obst:
        from . import abi
        from .scene import pack_obstacles, wall_polygon
        obst = np.zeros((N, max_obst), dtype=abi.OBST_DTYPE)
        obst_count = np.zeros(N, np.int32)
        for e in range(N):
            rec, _ = pack_obstacles([wall_polygon(*w) for w in walls[e]], max_obst)
            obst[e, :len(rec)] = rec
            obst_count[e] = len(rec)
        extra = {"obst": obst, "obst_count": obst_count}
    return {
        **extra,
        "hum_pv": hum_pv, "hum_gr": hum_gr, "hum_type": hum_type, "hum_count": np.full(N, H, np.int32),
        "stat": stat, "stat_count": stat_count, "rect": rect, "rect_count": rect_count,
        "rob_pv": rob_pv, "rob_gr": rob_gr, "rob_theta": np.full(N, np.pi / 2, np.float32),
        "time": np.zeros(N, np.float64),
    }


def load(sim, scenes, first=0):
    sim.load_episodes(first, scenes["hum_pv"], scenes["hum_gr"], scenes["hum_type"], scenes["hum_count"],
                      scenes["stat"], scenes["stat_count"], scenes["rect"], scenes["rect_count"],
                      scenes["rob_pv"], scenes["rob_gr"], scenes["rob_theta"], scenes["time"],
                      obst=scenes.get("obst"), obst_count=scenes.get("obst_count"))
