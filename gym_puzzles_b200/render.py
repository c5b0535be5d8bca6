"""Host render bridge: one env's state -> the polygon list the reference's viewer draws
(reference mrp00:528-592, mrp02:590-707), plus a dependency-free numpy rasteriser for `render(mode='rgb_array')`.
Rendering stays on the host and per env (north star: "rendering stays out of scope on the host"); nothing here runs in
the step path.  Geometry constants restate the reference's module constants (mrp00:38-67, 299-378; mrp02:39-67,
313-411)."""
import math

import numpy as np

from .abi import VARIANTS

COLORS = {"agent": (1.0, 1.0, 1.0), "block": (0.5, 0.5, 0.5), "wall": (0.2, 0.2, 0.2), "goal": (58 / 255, 153 / 255, 1.0),
          "background": (0.0, 0.0, 0.0)}   # mrp00:69-81, 540-545


def _box(hx, hy, cx=0.0, cy=0.0):
    return [(cx - hx, cy - hy), (cx + hx, cy - hy), (cx + hx, cy + hy), (cx - hx, cy + hy)]


def geometry(env_id):
    """Local-frame fixture polygons (body origin frame), block local centre, viewport and scale of one variant."""
    v = VARIANTS[env_id] if isinstance(env_id, str) else int(env_id)
    v2, heavy = v >= 2, bool(v & 1)
    if not v2:
        s = 1.0 if heavy else 2.0                                       # mrp00:303-309
        block = [_box(1 / s, 1 / s, 0.0, -1 / s), _box(3 / s, 1 / s, 0.0, 1 / s)]
        S = 2.0
        agent = [[(-0.5 / S, -1.5 / S), (0.5 / S, -1.5 / S), (1.5 / S, -0.5 / S), (1.5 / S, 0.5 / S), (0.5 / S, 1.5 / S),
                  (-0.5 / S, 1.5 / S), (-1.5 / S, 0.5 / S), (-1.5 / S, -0.5 / S)]]
        scale, vw, vh, wall = 30.0, 640, 480, 1.0
    else:
        block = [_box(0.1, 0.1, 0.0, -0.1), _box(0.3, 0.1, 0.0, 0.1)]   # mrp02:331-341
        agent = [[(-0.039, -0.095), (0.039, -0.095), (0.095, -0.039), (0.095, 0.039), (0.039, 0.095), (-0.039, 0.095),
                  (-0.095, 0.039), (-0.095, -0.039)], _box(0.005, 0.05, 0.06, 0.0), _box(0.005, 0.05, -0.06, 0.0)]
        scale, vw, vh, wall = 560.0, 1440, 810, 0.1
    W, H = vw / scale, vh / scale
    areas = [(p[1][0] - p[0][0]) * (p[2][1] - p[1][1]) for p in block]
    cys = [(p[0][1] + p[2][1]) / 2 for p in block]
    block_lc = (0.0, sum(a * c for a, c in zip(areas, cys)) / sum(areas))   # same density on both boxes
    walls = [_box(wall, H, 0.0, H / 2), _box(wall, H, W, H / 2), _box(W, wall, W / 2, 0.0), _box(W, wall, W / 2, H)]
    return dict(block=block, agent=agent, block_local_center=block_lc, walls=walls, scale=scale, viewport=(vw, vh), world=(W, H), v2=v2)


def _place(poly, cx, cy, angle, lc=(0.0, 0.0)):
    c, s = math.cos(angle), math.sin(angle)
    ox, oy = cx - (c * lc[0] - s * lc[1]), cy - (s * lc[0] + c * lc[1])     # body origin = worldCenter - R * localCenter
    return [(ox + c * x - s * y, oy + s * x + c * y) for x, y in poly]


def scene(handle, env_index=0):
    """-> {"polygons": [(kind, [(x, y), ...] world metres)], "goal": (x, y) metres, "world": (W, H), "viewport", "scale"}"""
    g = geometry(handle.variant)
    l = handle.layout
    w = handle.get_state(env_index, 1)[0]
    bodies = w[l.off_bodies:l.off_bodies + 6 * l.n_dyn_bodies].view(np.float32).reshape(l.n_dyn_bodies, 6)
    goal = np.ascontiguousarray(w[l.off_goal:l.off_goal + 4]).view(np.float64)
    polys = [("wall", p) for p in g["walls"]]
    bx, by, ba = (float(x) for x in bodies[0, :3])
    polys += [("block", _place(p, bx, by, ba, g["block_local_center"])) for p in g["block"]]
    for i in range(l.n_agents):
        ax, ay, aa = (float(x) for x in bodies[1 + i, :3])
        polys += [("agent", _place(p, ax, ay, aa)) for p in g["agent"]]
    # goal is stored in the env's reporting units: pixels (v0, mrp00:115-128) or normalised units (v2, mrp02:303-311)
    k = g["scale"] if not g["v2"] else g["scale"] / g["viewport"][0]
    return dict(polygons=polys, goal=(goal[0] / k, goal[1] / k), world=g["world"], viewport=g["viewport"], scale=g["scale"],
                goal_radius=(25.0 / g["scale"]) if not g["v2"] else 0.1 / k)


def rgb_array(handle, env_index=0, downsample=1):
    """uint8[H, W, 3] image of one env (origin bottom-left like the reference viewer, returned top-row-first)."""
    sc = scene(handle, env_index)
    vw, vh = sc["viewport"][0] // downsample, sc["viewport"][1] // downsample
    ppm = sc["scale"] / downsample
    img = np.zeros((vh, vw, 3), dtype=np.float32)
    img[...] = COLORS["background"]
    ys, xs = np.mgrid[0:vh, 0:vw]
    px, py = (xs + 0.5) / ppm, (ys + 0.5) / ppm

    def fill(poly, color):
        p = np.asarray(poly, dtype=np.float64)
        x0, x1 = max(int(p[:, 0].min() * ppm), 0), min(int(p[:, 0].max() * ppm) + 1, vw)
        y0, y1 = max(int(p[:, 1].min() * ppm), 0), min(int(p[:, 1].max() * ppm) + 1, vh)
        if x0 >= x1 or y0 >= y1:
            return
        X, Y = px[y0:y1, x0:x1], py[y0:y1, x0:x1]
        inside = np.ones(X.shape, dtype=bool)
        for k in range(len(p)):                    # convex, counter-clockwise: left of every edge
            ax, ay = p[k]
            bx, by = p[(k + 1) % len(p)]
            inside &= (bx - ax) * (Y - ay) - (by - ay) * (X - ax) >= 0
        img[y0:y1, x0:x1][inside] = color

    for kind, poly in sc["polygons"]:
        fill(poly, COLORS[kind])
    gx, gy, r = sc["goal"][0], sc["goal"][1], sc["goal_radius"]
    img[(px - gx) ** 2 + (py - gy) ** 2 <= r * r] = COLORS["goal"]
    return (img[::-1] * 255).astype(np.uint8)
