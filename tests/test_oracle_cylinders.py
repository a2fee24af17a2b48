"""CPU: the EXTENSION cylinder pairs of the restatement (sphere-cylinder, cylinder-box, cylinder-cylinder; mesh-cylinder is
in test_oracle_sdf.py). The reference's dispatch (CollisionDetect.cpp:45-73) only knows cylinder-plane and silently ignores
these pairs, so there is no reference behaviour to match: the checks are physical (resting contact at the right height,
small penetration, normals from body1 towards body0) and that the reference dispatch is untouched with the extension off."""
import numpy as np

from oracle_lib import Oracle, BOX, CYLINDER, PLANE, SPHERE

DT = 0.01
QX90 = [np.sin(np.pi / 4), 0, 0, np.cos(np.pi / 4)]          # local y (the cylinder axis) -> world z: lying on its side


def test_cylinder_pairs_are_ignored_without_the_extension():
    o = Oracle()
    o.add_body(1.0, BOX, [4, 0.4, 4], fixed=True)
    o.add_body(1.0, CYLINDER, [1.0, 0.5], x=[0, 0.69, 0])
    o.add_body(1.0, SPHERE, [0.3], x=[0, 1.45, 0])
    o.step(DT, 3)
    assert o.n_contacts == 0                                  # as in the reference: the wheels of the car touch only the plane


def test_standing_and_lying_cylinders_rest_on_a_box():
    o = Oracle(); o.set_extensions(1)
    o.add_body(1.0, BOX, [8, 0.4, 8], fixed=True)
    o.add_body(1.0, CYLINDER, [1.0, 0.5], x=[-2, 1.0, 0])                     # standing: rests at 0.2 + h/2
    o.add_body(1.0, CYLINDER, [1.0, 0.5], x=[2, 1.0, 0], q=QX90)              # lying: rests at 0.2 + r
    o.step(DT, 400)
    st = o.state()
    assert abs(st["x"][1, 1] - 0.7) < 5e-3, st["x"][1]
    assert abs(st["x"][2, 1] - 0.7) < 0.012, st["x"][2]       # the 16 samples of a ring leave r (1 - cos 11.25 deg) at worst
    assert np.abs(st["v"]).max() < 0.05 and np.abs(st["x"][1:, [0, 2]] - [[-2, 0], [2, 0]]).max() < 0.05
    c = o.contacts()
    assert c["phi"].min() > -0.012 and (c["body1"] == 0).all() and set(c["body0"].tolist()) == {1, 2}
    np.testing.assert_allclose(c["n"], np.tile([0, 1, 0], (o.n_contacts, 1)), atol=1e-5)


def test_sphere_on_cylinder_and_cylinder_on_cylinder():
    o = Oracle(); o.set_extensions(1)
    o.add_body(1.0, CYLINDER, [1.0, 1.0], x=[0, 0.5, 0], fixed=True)          # a fixed drum, top face at y = 1
    o.add_body(1.0, SPHERE, [0.3], x=[0.2, 1.6, 0.1])
    o.add_body(1.0, CYLINDER, [0.4, 0.3], x=[-0.4, 1.5, -0.2])                # a small cylinder standing on the drum
    o.step(DT, 300)
    st = o.state()
    assert abs(st["x"][1, 1] - 1.3) < 5e-3 and abs(st["x"][2, 1] - 1.2) < 5e-3, st["x"]
    c = o.contacts()
    pairs = set(zip(c["body0"].tolist(), c["body1"].tolist()))
    assert (1, 0) in pairs and ((2, 0) in pairs or (0, 2) in pairs)
    for b0, b1, n in zip(c["body0"], c["body1"], c["n"]):
        assert abs(np.linalg.norm(n) - 1) < 1e-5
        assert n[1] * (1 if b1 == 0 else -1) > 0.99           # from body1 towards body0
    # a sphere dropped beside the drum's edge rolls off instead of sticking (rim normal is oblique)
    p = Oracle(); p.set_extensions(1)
    p.add_body(1.0, CYLINDER, [1.0, 1.0], x=[0, 0.5, 0], fixed=True)
    p.add_body(1.0, SPHERE, [0.3], x=[1.1, 1.5, 0])
    p.step(DT, 120)
    assert p.state()["x"][1, 0] > 1.3 and p.state()["x"][1, 1] < 1.0


def test_car_wheels_touch_a_box_with_the_extension():
    """The reference car's wheels can touch nothing but the plane; with the extension a box on the road
    stops a wheel."""
    o = Oracle("car"); o.set_extensions(1)
    st = o.state()
    wheel = int(np.argmin(st["x"][1:5, 2])) + 1               # the front-most wheel (the car drives towards -z)
    o.add_body(1.0, BOX, [3.0, 1.0, 0.5], x=[st["x"][wheel, 0], 0.5, st["x"][wheel, 2] - 1.2], fixed=True)
    seen = False
    for _ in range(150):
        o.step(DT)
        c = o.contacts()
        seen = seen or bool(((c["body1"] == o.n_bodies - 1) | (c["body0"] == o.n_bodies - 1)).any())
    assert seen
    assert o.state()["x"][wheel, 2] > st["x"][wheel, 2] - 1.5  # it did not drive through
