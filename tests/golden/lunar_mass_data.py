"""Mass data of the two LunarLander polygons, evaluated the way b2PolygonShape::ComputeMass /
b2Body::ResetMassData do it (Box2D 2.3 lineage of Aether.Physics2D), in float32.

    fuselage: LANDER_POLY / SCALE, density 5      (LunarLanderEnv.cs:189, :238)
    leg:      box (0,0)-(LEG_W/S, LEG_H/S), density 1   (LunarLanderEnv.cs:262-267)

oracle/lunar.hpp and gym.net_b200/csrc/lunar_core.cuh hard-code the numbers printed here;
tests/test_oracle_lunar_cpu.py::test_mass_data_constants re-derives and compares them."""
import numpy as np

f = np.float32


def mass_data(verts, density):
    n = len(verts)
    V = [(f(x), f(y)) for x, y in verts]
    sx = f(0); sy = f(0)
    for x, y in V:
        sx = f(sx + x); sy = f(sy + y)
    inv = f(f(1) / f(n)); sx = f(sx * inv); sy = f(sy * inv)          # reference point s = vertex average
    area = f(0); cx = f(0); cy = f(0); inertia = f(0); k3 = f(f(1) / f(3))
    for i in range(n):
        e1 = (f(V[i][0] - sx), f(V[i][1] - sy)); j = (i + 1) % n; e2 = (f(V[j][0] - sx), f(V[j][1] - sy))
        D = f(f(e1[0] * e2[1]) - f(e1[1] * e2[0])); ta = f(f(0.5) * D); area = f(area + ta)
        cx = f(cx + f(f(ta * k3) * f(e1[0] + e2[0]))); cy = f(cy + f(f(ta * k3) * f(e1[1] + e2[1])))
        intx2 = f(f(f(e1[0] * e1[0]) + f(e2[0] * e1[0])) + f(e2[0] * e2[0]))
        inty2 = f(f(f(e1[1] * e1[1]) + f(e2[1] * e1[1])) + f(e2[1] * e2[1]))
        inertia = f(inertia + f(f(f(f(0.25) * k3) * D) * f(intx2 + inty2)))
    mass = f(f(density) * area)
    cx = f(cx * f(f(1) / area)); cy = f(cy * f(f(1) / area))
    centx = f(cx + sx); centy = f(cy + sy)
    io = f(f(density) * inertia)
    io = f(io + f(mass * f(f(f(centx * centx) + f(centy * centy)) - f(f(cx * cx) + f(cy * cy)))))   # about the body origin
    ic = f(io - f(mass * f(f(centx * centx) + f(centy * centy))))                                    # about the centre of mass
    return {"mass": float(mass), "inv_mass": float(f(f(1) / mass)), "inertia": float(ic), "inv_inertia": float(f(f(1) / ic)),
            "centroid": (float(centx), float(centy))}


SCALE = f(30)
FUSELAGE = [(float(f(x) / SCALE), float(f(y) / SCALE)) for x, y in [(17, -10), (17, 0), (14, 17), (-14, 17), (-17, 0), (-17, -10)]]
LEG_W, LEG_H = float(f(2) / SCALE), float(f(8) / SCALE)
LEG = [(LEG_W, 0.0), (LEG_W, LEG_H), (0.0, LEG_H), (0.0, 0.0)]


def constants():
    return {"fuselage": mass_data(FUSELAGE, 5.0), "leg": mass_data(LEG, 1.0)}


if __name__ == "__main__":
    for k, v in constants().items():
        print(k, v)
