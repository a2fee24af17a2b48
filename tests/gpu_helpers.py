"""Helpers shared by the GPU parity tests: build a device Context from oracle systems."""
import numpy as np

from slrbs_b200 import Context


def scene_arrays(o):
    """All upload arrays of one oracle system."""
    bp, st = o.body_params(), o.state()
    j = o.joints()
    return dict(x=st["x"], q=st["q"], v=st["v"], w=st["w"], mass=bp["mass"], ibody=bp["ibody"],
                ibodyinv=bp["ibodyinv"], fixed=bp["fixed"], geom_type=bp["geom_type"], geom=bp["geom"],
                jtype=j["type"], jb0=j["body0"], jb1=j["body1"], r0=j["r0"], q0=j["q0"], r1=j["r1"], q1=j["q1"])


def context_from_oracles(oracles, max_contacts=1024, device=0):
    """One scene per oracle system; all systems must have the same body/joint counts."""
    B = len(oracles)
    arrs = [scene_arrays(o) for o in oracles]
    nb, nj = oracles[0].n_bodies, oracles[0].n_joints
    assert all(o.n_bodies == nb and o.n_joints == nj for o in oracles)
    ctx = Context(B, nb, max(nj, 0), max_contacts, device=device)
    st = lambda k: np.stack([a[k] for a in arrs])
    ctx.upload_bodies(st("x"), st("q"), st("v"), st("w"), st("mass"), st("ibody"), st("ibodyinv"),
                      st("fixed"), st("geom_type"), st("geom"))
    if nj:
        ctx.upload_joints(st("jtype"), st("jb0"), st("jb1"), st("r0"), st("q0"), st("r1"), st("q1"))
    return ctx


def push_state(ctx, oracles):
    st = [o.state() for o in oracles]
    ctx.upload_state(*[np.stack([s[k] for s in st]) for k in ("x", "q", "v", "w")])
    if oracles[0].n_joints:
        ctx.set_joint_lambda(np.stack([o.joint_rows()["lam"] for o in oracles]))


def bits(a):
    return np.ascontiguousarray(np.asarray(a, np.float32)).view(np.uint32)


def assert_bit_equal(a, b, what=""):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    same = (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))
    if not same.all():
        idx = np.argwhere(~same)[:5]
        raise AssertionError(f"{what}: {(~same).sum()} of {same.size} values differ bitwise; first at {idx.tolist()}: "
                             f"{a[tuple(idx[0])]!r} vs {b[tuple(idx[0])]!r}")
