"""CPU: the EXTENSION contact pairs of the restatement (sphere-plane, box-plane, box-box). The
reference ignores these pairs, so there is no reference behaviour: the checks are that the reference dispatch is
untouched while the extension is off, and that the new contacts behave physically (resting contact, small
penetration, a 5-box stack that stays up)."""
import numpy as np

from oracle_lib import Oracle, BOX, PLANE, SPHERE

DT = 0.01


def test_extension_is_off_by_default_and_reference_scenes_do_not_change():
    a = Oracle("box_stack", k=3)
    a.set_extensions(0)
    a.step(DT, 5)
    assert a.n_contacts == 0                                 # box-box / box-plane ignored, as in the reference
    for scene in ("marble_box", "car", "sphere_on_box"):
        p, q = Oracle(scene), Oracle(scene)
        q.set_extensions(1)                                  # only adds pairs the reference ignores; these scenes have none
        p.step(DT, 40); q.step(DT, 40)
        if scene == "car":                                   # chassis box - ground plane is such a pair
            assert q.n_contacts >= p.n_contacts
        else:
            for k in "xqvw":
                np.testing.assert_array_equal(p.state()[k], q.state()[k])


def test_sphere_and_box_come_to_rest_on_a_plane():
    o = Oracle(); o.set_extensions(1)
    o.add_body(1.0, SPHERE, [0.5], x=[0, 2, 0])
    o.add_body(1.0, PLANE, [0, 0, 0, 0, 1, 0], fixed=True)
    q = np.array([0.2, 0.1, 0.3, 0.9]); q /= np.linalg.norm(q)
    o.add_body(2.0, BOX, [1, 0.6, 1.4], x=[3, 2, 0], q=q)
    o.step(DT, 500)
    st = o.state()
    assert abs(st["x"][0, 1] - 0.5) < 5e-3 and abs(st["x"][2, 1] - 0.3) < 5e-3
    assert np.abs(st["v"]).max() < 0.05
    c = o.contacts()
    assert o.n_contacts == 5 and c["phi"].min() > -5e-3
    assert (c["body0"] == [0, 2, 2, 2, 2]).all() and (c["body1"] == 1).all()
    np.testing.assert_allclose(c["n"], np.tile([0, 1, 0], (5, 1)), atol=1e-6)


def test_box_stack_stays_up():
    o = Oracle("box_stack", k=5)
    o.step(DT, 400)
    st = o.state()
    np.testing.assert_allclose(st["x"][:5, 1], 0.5 + np.arange(5), atol=0.01)
    assert np.abs(st["x"][:5, [0, 2]]).max() < 0.08 and o.contacts()["phi"].min() > -5e-3
    c = o.contacts()
    assert set(zip(c["body0"].tolist(), c["body1"].tolist())) == {(0, 1), (0, 5), (1, 2), (2, 3), (3, 4)}


def test_box_box_edge_contact_and_separation():
    o = Oracle(); o.set_extensions(1)
    o.add_body(1.0, BOX, [1, 1, 1], x=[0, 0, 0], fixed=True)
    s = np.sin(np.pi / 8); c = np.cos(np.pi / 8)
    # a box turned 45 deg about z whose lower-left face cuts the top-right edge of the fixed box
    o.add_body(1.0, BOX, [1, 1, 1], x=[0.9, 0.8, 0.0], q=[0, 0, s, c])
    o.add_body(1.0, BOX, [1, 1, 1], x=[5, 5, 5])
    o.stage_prepare(); o.stage_detect()
    cc = o.contacts()
    assert o.n_contacts >= 1 and (cc["body0"] == 0).all() and (cc["body1"] == 1).all()
    assert np.isfinite(cc["p"]).all() and abs(np.linalg.norm(cc["n"][0]) - 1) < 1e-5
    assert np.dot(cc["n"][0], o.state()["x"][0] - o.state()["x"][1]) > 0      # normal from body1 towards body0
