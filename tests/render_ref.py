"""numpy float32 restatement of gym.net_b200/csrc/render.cuh / lunar_render.cuh: what CartPoleEnv.Render
(CartPoleEnv.cs:69-135) and LunarLanderEnv.Render (LunarLanderEnv.cs:776-890) draw, sampled at pixel centres without
anti-aliasing.  TEST INFRASTRUCTURE (the checker of tests/test_gpu_round2.py::test_render_*)."""
import numpy as np

import oracle_lib as O

f32 = np.float32
CW, CH = f32(600.0), f32(400.0)


def _centres(width, height):
    xs = (np.arange(width, dtype=np.float32) + f32(0.5)) * f32(CW / f32(width))
    ys = (np.arange(height, dtype=np.float32) + f32(0.5)) * f32(CH / f32(height))
    return np.meshgrid(xs, ys)   # [h, w] each


def _seg_dist2(seg, px, py):
    ax, ay, bx, by = (f32(v) for v in seg)
    dx, dy = f32(bx - ax), f32(by - ay)
    len2 = f32(f32(dx * dx) + f32(dy * dy))
    if len2 > 0:
        t = (f32(px - ax) * dx + f32(py - ay) * dy).astype(np.float32) / len2
    else:
        t = np.zeros_like(px)
    t = np.clip(t, f32(0), f32(1)).astype(np.float32)
    qx = (ax + t * dx).astype(np.float32) - px
    qy = (ay + t * dy).astype(np.float32) - py
    return (qx * qx).astype(np.float32) + (qy * qy).astype(np.float32)


def cartpole(state, width=600, height=400):
    x, theta = f32(state[0]), f32(state[2])
    cx, cy = _centres(width, height)
    scale = f32(CW / f32(f32(2.4) * f32(2.0)))
    carty, polewidth, cartwidth, cartheight = f32(300), f32(10), f32(50), f32(30)
    poleheight = f32(scale * f32(f32(2.0) * f32(0.5)))
    center_x = f32(f32(x * scale) + f32(CW / f32(2)))
    img = np.full((height, width, 3), 255, np.uint8)
    img[(cy >= carty) & (cy < carty + f32(1))] = 0
    lx = (cx - center_x).astype(np.float32)
    img[(lx >= -cartwidth / 2) & (lx < cartwidth / 2) & (cy >= carty - cartheight / 2) & (cy < carty + cartheight / 2)] = 0
    sn, cs = O.sincosf(np.array([theta], np.float32)); sn, cs = f32(sn[0]), f32(cs[0])
    pivot_y = f32(carty - polewidth / 2)
    dx, dy = lx, (cy - pivot_y).astype(np.float32)
    ux = (cs * dx).astype(np.float32) + (sn * dy).astype(np.float32)
    uy = ((-sn * dx).astype(np.float32) + (cs * dy).astype(np.float32)).astype(np.float32) + pivot_y
    pole = (ux >= -polewidth / 2) & (ux < polewidth / 2) & (uy >= carty - poleheight) & (uy < carty)
    disc = ((dx * dx).astype(np.float32) + (dy * dy).astype(np.float32)) <= f32((polewidth / 2) * (polewidth / 2))
    img[pole | disc] = (204, 153, 102)
    return img


# lander geometry: gym.net_b200/csrc/lunar_core.cuh SHAPES (b2PolygonShape::Set order), local vertices and centroids
_V = [np.array([[0.5666666626930237, -0.3333333432674408], [0.5666666626930237, 0.0], [0.46666666865348816, 0.5666666626930237],
                [-0.46666666865348816, 0.5666666626930237], [-0.5666666626930237, 0.0], [-0.5666666626930237, -0.3333333432674408]], np.float32),
      np.array([[0.06666667014360428, 0.0], [0.06666667014360428, 0.2666666805744171], [0.0, 0.2666666805744171], [0.0, 0.0]], np.float32)]
_V.append(_V[1])
_C = [np.array([0.0, 0.10130719095468521], np.float32), np.array([0.03333333507180214, 0.13333334028720856], np.float32)]
_C.append(_C[1])


def lunar(state80, width=600, height=400):
    """state80: the [80] float32 state row of gymcuda_get_state (bodies at 7 i = c.x, c.y, a; terrain at 53..63)."""
    S, W, H = f32(30.0), f32(f32(600.0) / f32(30.0)), f32(f32(400.0) / f32(30.0))
    st = np.asarray(state80, np.float32)
    cx, cy = _centres(width, height)
    sx, sy = f32(CW / f32(width)), f32(CH / f32(height))
    half = f32(f32(0.5) * max(sx, sy)); h2 = f32(half * half)
    img = np.zeros((height, width, 3), np.uint8)
    for bi in range(3):
        bcx, bcy, ba = st[7 * bi], st[7 * bi + 1], st[7 * bi + 2]
        sn, cs = O.sincosf(np.array([ba], np.float32)); sn, cs = f32(sn[0]), f32(cs[0])
        px = f32(bcx - f32(f32(cs * _C[bi][0]) - f32(sn * _C[bi][1]))); py = f32(bcy - f32(f32(sn * _C[bi][0]) + f32(cs * _C[bi][1])))
        V = _V[bi]; cnt = len(V)
        pts = []
        for k in range(cnt):
            x = f32(f32(f32(cs * V[k][0]) - f32(sn * V[k][1])) + px); y = f32(f32(f32(sn * V[k][0]) + f32(cs * V[k][1])) + py)
            pts.append((f32(x * S), f32(CH - f32(y * S))))
        for k in range(cnt):
            a, b = pts[k], pts[(k + 1) % cnt]
            img[_seg_dist2((a[0], a[1], b[0], b[1]), cx, cy) <= h2] = (128, 102, 230)
    ty = st[53:64]
    cw = f32(W / f32(10.0))
    moon = [(f32(f32(cw * f32(i)) * S), f32(CH - f32(ty[i] * S)), f32(f32(cw * f32(i + 1)) * S), f32(CH - f32(ty[i + 1] * S))) for i in range(10)]
    moon.append((f32(0), CH, f32(W * S), CH))
    cwp = f32(cw * S)
    idx = np.clip((cx / cwp).astype(np.int32), 0, 9)
    ax = np.array([m[0] for m in moon[:10]], np.float32)[idx]; ay = np.array([m[1] for m in moon[:10]], np.float32)[idx]
    bx = np.array([m[2] for m in moon[:10]], np.float32)[idx]; by = np.array([m[3] for m in moon[:10]], np.float32)[idx]
    t = ((cx - ax).astype(np.float32) / (bx - ax).astype(np.float32)).astype(np.float32)
    yline = (ay + (t * (by - ay).astype(np.float32)).astype(np.float32)).astype(np.float32)
    img[cy >= yline] = 255
    for m in moon:
        img[_seg_dist2(m, cx, cy) <= h2] = (255, 0, 0)
    for f in range(2):
        x1 = f32(f32(cw * f32(4.0 if f == 0 else 6.0)) * S)
        y1 = f32(CH - f32(f32(H / f32(4.0)) * S)); y2 = f32(y1 - f32(50.0))
        img[_seg_dist2((x1, y1, x1, y2), cx, cy) <= h2] = 255
        for seg in ((x1, y2, x1, f32(y2 + f32(10))), (x1, f32(y2 + f32(10)), f32(x1 + f32(25)), f32(y2 + f32(5))), (f32(x1 + f32(25)), f32(y2 + f32(5)), x1, y2)):
            img[_seg_dist2(seg, cx, cy) <= h2] = (204, 204, 0)
    return img
