"""Placement rules of the reference's scene generator, as checks on a batch of generated scenes (ebc_state layout).
Every rule cites the reference code it restates; the checks are applied both to the host generator of ebc/synth.py
(CPU suite) and to the output of the device generator ebc_generate (GPU suite)."""
import numpy as np


def check_scene_rules(shape, sc, max_fallback_frac=2e-3):
    """-> dict of counters; raises AssertionError on a violated rule."""
    N = len(sc["hum_count"])
    H = shape.H
    R, hw, dd = shape.circle_radius, shape.square_width / 2.0, shape.discomfort_dist
    pv, gr = sc["hum_pv"].astype(np.float64), sc["hum_gr"].astype(np.float64)
    p, g, vpref, rad = pv[:, :, :2], gr[:, :, :2], gr[:, :, 2], gr[:, :, 3]
    types = np.concatenate([np.full(c, t, np.uint8) for t, c, _, _ in shape.types])
    assert (sc["hum_count"] == H).all()
    assert (sc["hum_type"][:, :H] == types[None]).all()              # list order: adults, bicycles, children (env.py:392)
    assert (pv[:, :, 2:] == 0).all()                                  # agents start at rest (agent.set(px, py, gx, gy, 0, 0, 0))
    # scene_generator.py:292-328 / agent.sample_random_attributes: v_pref and radius inside the configured ranges
    h0 = 0
    for t, c, (v0, v1), (r0, r1) in shape.types:
        sl = slice(h0, h0 + c)
        assert (vpref[:, sl] >= np.float32(v0) - 1e-6).all() and (vpref[:, sl] <= np.float32(v1) + 1e-6).all()
        assert (rad[:, sl] >= np.float32(r0) - 1e-6).all() and (rad[:, sl] <= np.float32(r1) + 1e-6).all()
        h0 += c
    # robot: env.py:128-140
    assert np.allclose(sc["rob_pv"], np.array([0, -R, 0, 0])[None]) and np.allclose(sc["rob_gr"][:, :2], np.array([0, R])[None])
    assert (sc["time"] == 0).all() and np.allclose(sc["rob_theta"], np.pi / 2)
    on_circle = np.abs(np.hypot(p[..., 0], p[..., 1]) - R) < 1e-5
    opposite = np.abs(g + p).max(-1) < 1e-5
    is_circle = on_circle & opposite                                  # :593-618: start on the circle, goal = -start
    static = np.abs(g - p).max(-1) == 0                               # :463-487: static adults have goal == start
    on_side = (np.abs(np.abs(p[..., 0]) - hw) < 1e-5) | (np.abs(np.abs(p[..., 1]) - hw) < 1e-5)
    goal_opposite = ((np.abs(p[..., 1] - hw) < 1e-5) & (np.abs(g[..., 1] + hw) < 1e-5)) | \
                    ((np.abs(p[..., 1] + hw) < 1e-5) & (np.abs(g[..., 1] - hw) < 1e-5)) | \
                    ((np.abs(p[..., 0] + hw) < 1e-5) & (np.abs(g[..., 0] - hw) < 1e-5)) | \
                    ((np.abs(p[..., 0] - hw) < 1e-5) & (np.abs(g[..., 0] + hw) < 1e-5))
    is_square = on_side & goal_opposite & ~is_circle                  # :621-712: start on a side, goal on the opposite side
    if shape.rule == "circle_crossing":
        assert is_circle.all()
    elif shape.rule == "square_crossing":
        assert is_square.all()
    elif shape.rule == "mixed":                                       # :492-501: first half circle, second half square
        assert is_circle[:, :H // 2].all() and is_square[:, H // 2:].all()
    elif shape.rule == "one_static":                                  # :583-589: two adults standing still, fixed places
        assert H == 2 and static.all()
        assert (p[:, 0] == np.array([-2.0, -8.0])).all() and (p[:, 1] == np.array([-3.0, -8.0])).all()
    elif shape.rule == "mixed_20":                                    # :577-582
        n_static = static.sum(1)
        assert ((n_static >= 0) & (n_static <= 19)).all() and len(np.unique(n_static)) > 10      # randint(20)
        for e in range(N):
            s, d = int(n_static[e]), H - int(n_static[e])
            assert static[e, :s].all()
            if s:
                assert tuple(p[e, 0]) == (-0.5, -2.5)                 # :463-466
                assert (np.abs(p[e, 1:s, 0]) <= 3.0 + 1e-6).all() and (np.abs(p[e, 1:s, 1]) <= 4.0 + 1e-6).all()   # 6 x 8 box
            assert is_circle[e, s:s + d // 2].all() and is_square[e, s + d // 2:].all()
    # minimum separation (:608-616, :683-693, :472-483): a start keeps r_a + r_b + discomfort_dist from the robot's start
    # and from every earlier agent OF THE SAME LIST; circle-crossing starts also from their goals and the robot's goal.
    # The reference retries 100000 times; the generators here 64 and then keep the last draw -> a counted exception.
    viol = 0
    robot = np.array([0.0, -R])
    goal = np.array([0.0, R])
    for h in range(H):
        md_r = rad[:, h] + shape.robot_radius + dd
        bad = np.hypot(*(p[:, h] - robot).T) < md_r - 1e-6
        bad |= is_circle[:, h] & (np.hypot(*(p[:, h] - goal).T) < md_r - 1e-6)
        fixed = static[:, h] & ((h == 0) | (shape.rule == "one_static"))      # placed without sampling
        for j in range(h):
            if types[j] != types[h]:
                continue
            md = rad[:, h] + rad[:, j] + dd
            bad |= np.hypot(*(p[:, h] - p[:, j]).T) < md - 1e-6
            bad |= is_circle[:, h] & (np.hypot(*(p[:, h] - g[:, j]).T) < md - 1e-6)
        viol += int((bad & ~fixed).sum())
    assert viol <= max_fallback_frac * N * H + 1, viol
    # walls: :194-243 -- whole-metre length in [min, max], thickness 1 m, centre on the 0.1 m grid, clear of the robot's
    # start and goal by robot.radius + discomfort_dist in the axis-aligned sense; grid + discs: synth.wall_geometry
    G = int(round(shape.map_size_m / shape.map_resolution))
    assert (sc["rect_count"] == shape.num_walls).all()
    near = 0
    for w in range(shape.num_walls):
        r = sc["rect"][:, w].astype(np.int64)
        assert (r[:, 0] >= 1).all() and (r[:, 2] <= G).all() and (r[:, 1] >= 1).all() and (r[:, 3] <= G).all()
        wx, wy = r[:, 2] - r[:, 0], r[:, 3] - r[:, 1]
        thick = int(round(1.0 / shape.map_resolution))
        inside = (r[:, 0] > 1) & (r[:, 2] < G) & (r[:, 1] > 1) & (r[:, 3] < G)      # not clipped by the map border
        lens = np.maximum(wx, wy)[inside] * shape.map_resolution
        assert (np.minimum(wx, wy)[inside] == thick).all()
        assert (lens >= shape.wall_len[0] - 1e-9).all() and (lens <= shape.wall_len[1] + 1e-9).all()
        assert np.allclose(lens, np.rint(lens))
        # clearance from the robot's start / goal (only decidable for unclipped walls)
        cx = (r[:, 0] + r[:, 2]) / 2.0 * shape.map_resolution - shape.map_size_m / 2.0
        cy = (r[:, 1] + r[:, 3]) / 2.0 * shape.map_resolution - shape.map_size_m / 2.0
        xd, yd = wx * shape.map_resolution, wy * shape.map_resolution
        clear = shape.robot_radius + dd
        for gy in (-R, R):
            near += int((inside & (np.abs(cx) < xd / 2 + clear - 1e-9) & (np.abs(cy - gy) < yd / 2 + clear - 1e-9)).sum())
    assert near <= max_fallback_frac * N * max(shape.num_walls, 1) + 1, near
    # static discs: radius = half thickness * sqrt(2) (:380-422), never more than the longest wall decomposes into
    S = sc["stat_count"]
    assert (S >= shape.num_walls).all() if shape.wall_len[0] >= 1 else True
    assert (S <= shape.num_walls * shape.discs_per_wall).all()
    for k in range(sc["stat"].shape[1]):
        live = k < S
        assert np.allclose(sc["stat"][live, k, 2], 0.5 * np.sqrt(2.0), atol=1e-6)
        assert (sc["stat"][~live, k] == 0).all()
    return {"fallback_placements": viol, "walls_near_robot": near}
