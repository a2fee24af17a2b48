"""ctypes binding for the CPU oracle (oracle/libslrbs_oracle.so) and, when it has been built, for
oracle/_ref/libslrbs_ref.so = the reference's own sources compiled against the Eigen API shim
(oracle/ref_shim).  Both export the same orc_* entry points; `Oracle(scene)` is the C restatement,
`Reference(scene)` the reference's code.

TEST INFRASTRUCTURE ONLY.  The product (slrbs_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")
_LIB_PATH = os.path.join(_ORACLE_DIR, "libslrbs_oracle.so")
_REF_PATH = os.path.join(_ORACLE_DIR, "_ref", "libslrbs_ref.so")

SPHERE, BOX, PLANE, CYLINDER = 0, 1, 2, 3
SDF = 4  # extension: mesh as signed-distance grid
CONTACT, SPHERICAL, HINGE = 0, 1, 2

_lib = None
_ref = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", _ORACLE_DIR])


def ref_available():
    """oracle/_ref is built here (where /root/reference exists) and travels prebuilt to the GPU box."""
    return os.path.exists(_REF_PATH)


class _Optional:
    """Symbols the reference driver does not export (per-stage entry points) resolve to a stub."""

    def __init__(self, L):
        self._L = L

    def __getattr__(self, name):
        try:
            return getattr(self._L, name)
        except AttributeError:
            class _Missing:
                argtypes = None
                restype = None

                def __call__(self, *a):
                    raise NotImplementedError(f"{name} is not exported by oracle/_ref")
            m = _Missing()
            setattr(self, name, m)
            return m


def lib(impl="oracle"):
    global _lib, _ref
    if impl == "ref":
        if _ref is None:
            if not ref_available():
                raise FileNotFoundError(_REF_PATH + " (make -C oracle ref)")
            _ref = _bind(_Optional(C.CDLL(_REF_PATH)))
        return _ref
    if _lib is not None:
        return _lib
    src = os.path.join(_ORACLE_DIR, "slrbs_oracle.c")
    if (not os.path.exists(_LIB_PATH)) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)
    ):
        build_oracle()
    _lib = _bind(C.CDLL(_LIB_PATH))
    return _lib


def _bind(L):
    fp = C.POINTER(C.c_float)
    ip = C.POINTER(C.c_int)
    vp = C.c_void_p
    L.orc_create.restype = vp
    L.orc_destroy.argtypes = [vp]
    L.orc_clear.argtypes = [vp]
    L.orc_add_body.argtypes = [vp, C.c_float, C.c_int, fp, C.c_int, fp, fp, fp, fp]
    L.orc_add_body.restype = C.c_int
    L.orc_add_hinge.argtypes = [vp, C.c_int, C.c_int, fp, fp, fp, fp]
    L.orc_add_hinge.restype = C.c_int
    L.orc_add_spherical.argtypes = [vp, C.c_int, C.c_int, fp, fp]
    L.orc_add_spherical.restype = C.c_int
    L.orc_set_params.argtypes = [vp, C.c_float, C.c_int, C.c_int, C.c_int]
    L.orc_step.argtypes = [vp, C.c_float]
    for name in ("orc_stage_prepare", "orc_stage_detect", "orc_stage_jacobians",
                 "orc_save_state", "orc_restore_state"):
        getattr(L, name).argtypes = [vp]
    for name in ("orc_stage_solve", "orc_stage_scatter", "orc_stage_integrate"):
        getattr(L, name).argtypes = [vp, C.c_float]
    L.orc_add_force_at_pos.argtypes = [vp, C.c_int, fp, fp]
    L.orc_add_force.argtypes = [vp, C.c_int, fp, fp]
    for name in ("orc_num_bodies", "orc_num_joints", "orc_num_contacts", "orc_num_joint_rows"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = C.c_int
    L.orc_pairs_tested.argtypes = [vp]
    L.orc_pairs_tested.restype = C.c_longlong
    L.orc_get_state.argtypes = [vp, fp, fp, fp, fp]
    L.orc_set_state.argtypes = [vp, fp, fp, fp, fp]
    L.orc_get_body_params.argtypes = [vp, fp, fp, fp, ip, ip, fp]
    L.orc_get_body_dyn.argtypes = [vp, fp, fp, fp, fp, fp, fp]
    L.orc_get_contacts.argtypes = [vp, ip, ip, fp, fp, fp, fp, fp, fp]
    L.orc_get_contact_rows.argtypes = [vp, fp, fp, fp, fp]
    L.orc_get_joints.argtypes = [vp, ip, ip, ip, fp, fp, fp, fp]
    L.orc_get_joint_rows.argtypes = [vp, fp, fp, fp, fp, fp, fp]
    L.orc_set_joint_lambda.argtypes = [vp, fp]
    L.orc_get_solver_blocks.argtypes = [vp, C.c_float, fp, fp, fp, fp]
    for name in ("orc_scene_marble_box", "orc_scene_sphere_on_box", "orc_scene_swinging_boxes",
                 "orc_scene_cylinder_on_plane", "orc_scene_car"):
        getattr(L, name).argtypes = [vp]
    L.orc_set_extensions.argtypes = [vp, C.c_int]
    L.orc_blcp_residual.argtypes = [vp, C.c_float]
    L.orc_add_sdf_shape.argtypes = [vp, C.c_int, C.c_int, C.c_int, fp, C.c_float, fp, C.c_int, fp]
    L.orc_add_sdf_shape.restype = C.c_int
    L.orc_build_sdf.argtypes = [C.c_int, fp, C.c_int, ip, C.c_int, C.c_int, C.c_int, fp, C.c_float, fp]
    L.orc_load_obj.argtypes = [C.c_char_p, fp, C.c_int, ip, C.c_int, ip, ip]
    L.orc_load_obj.restype = C.c_int
    L.orc_blcp_residual.restype = C.c_float
    L.orc_scene_box_stack.argtypes = [vp, C.c_int]
    L.orc_scene_sphere_stack.argtypes = [vp, C.c_int]
    L.orc_scene_marble_box_n.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.orc_time_steps.argtypes = [vp, C.c_float, C.c_int]
    L.orc_time_steps.restype = C.c_double
    return L


def _f(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


def _fa(v, n):
    if v is None:
        return None
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(n))
    return a


SCENES = {
    "marble_box": "orc_scene_marble_box",
    "sphere_on_box": "orc_scene_sphere_on_box",
    "swinging_boxes": "orc_scene_swinging_boxes",
    "cylinder_on_plane": "orc_scene_cylinder_on_plane",
    "car": "orc_scene_car",
}


class Oracle:
    """One RigidBodySystem of the restated reference."""

    impl = "oracle"

    def __init__(self, scene=None, **kw):
        self.L = lib(self.impl)
        self.h = self.L.orc_create()
        if scene is not None:
            self.load_scene(scene, **kw)

    def __del__(self):
        try:
            if self.h:
                self.L.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- scenes ------------------------------------------------------
    def load_scene(self, scene, **kw):
        if scene in SCENES:
            getattr(self.L, SCENES[scene])(self.h)
        elif scene == "box_stack":
            self.L.orc_scene_box_stack(self.h, int(kw.get("k", 5)))
        elif scene == "sphere_stack":
            self.L.orc_scene_sphere_stack(self.h, int(kw.get("k", 8)))
        elif scene == "marble_box_n":
            self.L.orc_scene_marble_box_n(self.h, int(kw["nx"]), int(kw["nz"]), int(kw["ny"]))
        else:
            raise ValueError(scene)

    # -- setup -------------------------------------------------------
    def add_body(self, mass, geom_type, geom, fixed=False, x=None, q=None, v=None, w=None):
        g = np.zeros(6, np.float32)
        geom = np.asarray(geom, np.float32).ravel()
        g[: geom.size] = geom
        xs, qs, vs, ws = _fa(x, 3), _fa(q, 4), _fa(v, 3), _fa(w, 3)
        return self.L.orc_add_body(self.h, float(mass), int(geom_type), _f(g), int(bool(fixed)),
                                   _f(xs), _f(qs), _f(vs), _f(ws))

    def add_hinge(self, b0, b1, r0, q0, r1, q1):
        a = [_fa(r0, 3), _fa(q0, 4), _fa(r1, 3), _fa(q1, 4)]
        return self.L.orc_add_hinge(self.h, b0, b1, *[_f(t) for t in a])

    def add_spherical(self, b0, b1, r0, r1):
        a = [_fa(r0, 3), _fa(r1, 3)]
        return self.L.orc_add_spherical(self.h, b0, b1, *[_f(t) for t in a])

    def set_params(self, mu=0.8, solver_id=0, solver_iter=10, collisions=True):
        self.L.orc_set_params(self.h, float(mu), int(solver_id), int(solver_iter), int(bool(collisions)))

    def add_sdf_shape(self, shape):
        """shape: dict(dims=(nx,ny,nz), origin, cell, grid, verts) as mesh_lib.make_shape returns."""
        nx, ny, nz = shape["dims"]
        g = np.ascontiguousarray(shape["grid"], np.float32); v = np.ascontiguousarray(shape["verts"], np.float32)
        o = np.asarray(shape["origin"], np.float32)
        return self.L.orc_add_sdf_shape(self.h, nx, ny, nz, _f(o), float(shape["cell"]), _f(g), len(v), _f(v))

    def set_extensions(self, flags=1):
        """Extension pairs without a reference counterpart (sphere-plane, box-plane, box-box)."""
        self.L.orc_set_extensions(self.h, int(flags))

    # -- stepping ----------------------------------------------------
    def step(self, dt=0.01, n=1):
        for _ in range(n):
            self.L.orc_step(self.h, float(dt))

    def time_steps(self, dt, n):
        return self.L.orc_time_steps(self.h, float(dt), int(n))

    def stage_prepare(self): self.L.orc_stage_prepare(self.h)
    def stage_detect(self): self.L.orc_stage_detect(self.h)
    def stage_jacobians(self): self.L.orc_stage_jacobians(self.h)
    def stage_solve(self, dt): self.L.orc_stage_solve(self.h, float(dt))
    def blcp_residual(self, dt): return float(self.L.orc_blcp_residual(self.h, float(dt)))
    def stage_scatter(self, dt): self.L.orc_stage_scatter(self.h, float(dt))
    def stage_integrate(self, dt): self.L.orc_stage_integrate(self.h, float(dt))
    def save_state(self): self.L.orc_save_state(self.h)
    def restore_state(self): self.L.orc_restore_state(self.h)

    def add_force(self, body, f=None, tau=None):
        a, b = _fa(f, 3), _fa(tau, 3)
        self.L.orc_add_force(self.h, int(body), _f(a), _f(b))

    def add_force_at_pos(self, body, pos, force):
        a, b = _fa(pos, 3), _fa(force, 3)
        self.L.orc_add_force_at_pos(self.h, int(body), _f(a), _f(b))

    # -- sizes -------------------------------------------------------
    @property
    def n_bodies(self): return self.L.orc_num_bodies(self.h)
    @property
    def n_joints(self): return self.L.orc_num_joints(self.h)
    @property
    def n_contacts(self): return self.L.orc_num_contacts(self.h)
    @property
    def n_joint_rows(self): return self.L.orc_num_joint_rows(self.h)
    @property
    def pairs_tested(self): return self.L.orc_pairs_tested(self.h)

    # -- getters -----------------------------------------------------
    def state(self):
        n = self.n_bodies
        x = np.zeros((n, 3), np.float32); q = np.zeros((n, 4), np.float32)
        v = np.zeros((n, 3), np.float32); w = np.zeros((n, 3), np.float32)
        self.L.orc_get_state(self.h, _f(x), _f(q), _f(v), _f(w))
        return dict(x=x, q=q, v=v, w=w)

    def set_state(self, x=None, q=None, v=None, w=None):
        n = self.n_bodies
        a = [None if t is None else np.ascontiguousarray(np.asarray(t, np.float32).reshape(n, k))
             for t, k in ((x, 3), (q, 4), (v, 3), (w, 3))]
        self.L.orc_set_state(self.h, *[_f(t) for t in a])

    def body_params(self):
        n = self.n_bodies
        mass = np.zeros(n, np.float32); ib = np.zeros((n, 9), np.float32); ibi = np.zeros((n, 9), np.float32)
        fixed = np.zeros(n, np.int32); gt = np.zeros(n, np.int32); g = np.zeros((n, 6), np.float32)
        self.L.orc_get_body_params(self.h, _f(mass), _f(ib), _f(ibi), _i(fixed), _i(gt), _f(g))
        return dict(mass=mass, ibody=ib, ibodyinv=ibi, fixed=fixed, geom_type=gt, geom=g)

    def body_dyn(self):
        n = self.n_bodies
        out = {k: np.zeros((n, 3), np.float32) for k in ("f", "tau", "fc", "tauc")}
        out["I"] = np.zeros((n, 9), np.float32); out["Iinv"] = np.zeros((n, 9), np.float32)
        self.L.orc_get_body_dyn(self.h, _f(out["f"]), _f(out["tau"]), _f(out["fc"]), _f(out["tauc"]),
                                _f(out["I"]), _f(out["Iinv"]))
        return out

    def contacts(self):
        m = self.n_contacts
        b0 = np.zeros(m, np.int32); b1 = np.zeros(m, np.int32)
        p = np.zeros((m, 3), np.float32); n = np.zeros((m, 3), np.float32)
        t = np.zeros((m, 3), np.float32); b = np.zeros((m, 3), np.float32)
        phi = np.zeros(m, np.float32); lam = np.zeros((m, 3), np.float32)
        self.L.orc_get_contacts(self.h, _i(b0), _i(b1), _f(p), _f(n), _f(t), _f(b), _f(phi), _f(lam))
        return dict(body0=b0, body1=b1, p=p, n=n, t=t, b=b, phi=phi, lam=lam)

    def contact_rows(self):
        m = self.n_contacts
        out = {k: np.zeros((m, 3, 6), np.float32) for k in ("J0", "J1", "J0Minv", "J1Minv")}
        self.L.orc_get_contact_rows(self.h, _f(out["J0"]), _f(out["J1"]), _f(out["J0Minv"]), _f(out["J1Minv"]))
        return out

    def joints(self):
        m = self.n_joints
        ty = np.zeros(m, np.int32); b0 = np.zeros(m, np.int32); b1 = np.zeros(m, np.int32)
        r0 = np.zeros((m, 3), np.float32); q0 = np.zeros((m, 4), np.float32)
        r1 = np.zeros((m, 3), np.float32); q1 = np.zeros((m, 4), np.float32)
        self.L.orc_get_joints(self.h, _i(ty), _i(b0), _i(b1), _f(r0), _f(q0), _f(r1), _f(q1))
        return dict(type=ty, body0=b0, body1=b1, r0=r0, q0=q0, r1=r1, q1=q1)

    def joint_rows(self):
        m = self.n_joints
        out = {k: np.zeros((m, 5, 6), np.float32) for k in ("J0", "J1", "J0Minv", "J1Minv")}
        out["phi"] = np.zeros((m, 5), np.float32); out["lam"] = np.zeros((m, 5), np.float32)
        self.L.orc_get_joint_rows(self.h, _f(out["J0"]), _f(out["J1"]), _f(out["J0Minv"]), _f(out["J1Minv"]),
                                  _f(out["phi"]), _f(out["lam"]))
        return out

    def solver_blocks(self, dt):
        m, nj = self.n_contacts, self.n_joints
        A = np.zeros((m, 3, 3), np.float32); rhs = np.zeros((m, 3), np.float32)
        Aj = np.zeros((nj, 5, 5), np.float32); rhsj = np.zeros((nj, 5), np.float32)
        self.L.orc_get_solver_blocks(self.h, float(dt), _f(A), _f(rhs), _f(Aj), _f(rhsj))
        return dict(A=A, rhs=rhs, Aj=Aj, rhsj=rhsj)

    def set_joint_lambda(self, lam):
        a = np.ascontiguousarray(np.asarray(lam, np.float32).reshape(self.n_joints, 5))
        self.L.orc_set_joint_lambda(self.h, _f(a))


class Reference(Oracle):
    """Same interface, backed by the reference's own sources (oracle/_ref)."""
    impl = "ref"
