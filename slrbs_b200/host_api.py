"""ctypes binding of include/slrbs_host.h (libslrbs_host.so): the C++ mirror of the reference's
RigidBodySystem / Scenarios API. Scene building needs no GPU; stepping needs a B200."""
import ctypes as C
import os

import numpy as np

from .capi import SlrbsError, load_library

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SYMBOLS = [
    "slrbs_host_last_error", "slrbs_host_scene_create", "slrbs_host_scene_destroy", "slrbs_host_num_bodies",
    "slrbs_host_num_joints", "slrbs_host_num_contacts", "slrbs_host_export_bodies", "slrbs_host_export_joints",
    "slrbs_host_set_solver", "slrbs_host_set_prestep_force", "slrbs_host_set_state", "slrbs_host_step",
    "slrbs_host_get_state", "slrbs_host_get_contacts", "slrbs_host_get_joint_lambda", "slrbs_host_save",
    "slrbs_host_restore", "slrbs_host_use_host_solver", "slrbs_host_batch_create", "slrbs_host_batch_destroy",
    "slrbs_host_batch_set_solver", "slrbs_host_batch_step", "slrbs_host_batch_sync", "slrbs_host_batch_set_state",
    "slrbs_host_batch_get_state", "slrbs_host_batch_set_forces", "slrbs_host_batch_save", "slrbs_host_batch_reset",
    "slrbs_host_batch_contact_counts", "slrbs_host_batch_read_scene",
]


def host_lib_path():
    return os.path.join(_HERE, "libslrbs_host.so")


def load_host_library():
    global _LIB
    if _LIB is not None:
        return _LIB
    load_library()      # libslrbs_b200.so first (also resolved through the rpath)
    p = host_lib_path()
    if not os.path.exists(p):
        raise SlrbsError(f"{p} is missing: build it with `make -C slrbs_b200/host`")
    L = C.CDLL(p)
    fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_void_p
    L.slrbs_host_last_error.restype = C.c_char_p
    L.slrbs_host_scene_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
    L.slrbs_host_scene_create.restype = vp
    L.slrbs_host_scene_destroy.argtypes = [vp]
    L.slrbs_host_scene_destroy.restype = None
    for n in ("slrbs_host_num_bodies", "slrbs_host_num_joints", "slrbs_host_num_contacts"):
        getattr(L, n).argtypes = [vp]
    L.slrbs_host_export_bodies.argtypes = [vp, fp, fp, fp, fp, fp, fp, fp, ip, ip, fp]
    L.slrbs_host_export_joints.argtypes = [vp, ip, ip, ip, fp, fp, fp, fp, fp]
    L.slrbs_host_set_solver.argtypes = [vp, C.c_int, C.c_int, C.c_float, C.c_int]
    L.slrbs_host_set_prestep_force.argtypes = [vp, C.c_int, fp, fp]
    L.slrbs_host_set_state.argtypes = [vp, fp, fp, fp, fp]
    L.slrbs_host_step.argtypes = [vp, C.c_float, C.c_int, C.c_int]
    L.slrbs_host_get_state.argtypes = [vp, fp, fp, fp, fp]
    L.slrbs_host_get_contacts.argtypes = [vp, ip, ip, fp, fp, fp, fp]
    L.slrbs_host_get_joint_lambda.argtypes = [vp, fp]
    L.slrbs_host_save.argtypes = [vp]
    L.slrbs_host_restore.argtypes = [vp]
    for n in ("slrbs_host_export_bodies", "slrbs_host_export_joints", "slrbs_host_set_solver", "slrbs_host_set_prestep_force",
              "slrbs_host_set_state", "slrbs_host_get_state", "slrbs_host_get_contacts", "slrbs_host_get_joint_lambda",
              "slrbs_host_save", "slrbs_host_restore"):
        getattr(L, n).restype = None
    L.slrbs_host_use_host_solver.argtypes = [vp, C.c_int]
    L.slrbs_host_use_host_solver.restype = None
    L.slrbs_host_batch_create.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.slrbs_host_batch_create.restype = vp
    L.slrbs_host_batch_destroy.argtypes = [vp]
    L.slrbs_host_batch_destroy.restype = None
    L.slrbs_host_batch_set_solver.argtypes = [vp, C.c_int, C.c_int, C.c_float, C.c_int]
    L.slrbs_host_batch_step.argtypes = [vp, C.c_float, C.c_int]
    L.slrbs_host_batch_sync.argtypes = [vp]
    L.slrbs_host_batch_set_state.argtypes = [vp, C.c_int, C.c_int, fp, fp, fp, fp]
    L.slrbs_host_batch_get_state.argtypes = [vp, C.c_int, C.c_int, fp, fp, fp, fp]
    L.slrbs_host_batch_set_forces.argtypes = [vp, C.c_int, C.c_int, fp, fp, C.c_int]
    L.slrbs_host_batch_save.argtypes = [vp]
    L.slrbs_host_batch_reset.argtypes = [vp, C.c_int, C.c_int]
    L.slrbs_host_batch_contact_counts.argtypes = [vp, C.c_int, C.c_int, ip]
    L.slrbs_host_batch_read_scene.argtypes = [vp, C.c_int, vp]
    _LIB = L
    return L


def _f(a):
    return None if a is None else C.cast(a.ctypes.data, C.POINTER(C.c_float))


def _i(a):
    return None if a is None else C.cast(a.ctypes.data, C.POINTER(C.c_int32))


class HostSystem:
    """One RigidBodySystem built by a Scenarios::create* function of the C++ mirror."""

    def __init__(self, scene, a=0, b=0, c=0):
        self.L = load_host_library()
        self.h = self.L.slrbs_host_scene_create(scene.encode(), int(a), int(b), int(c))
        if not self.h:
            raise SlrbsError(self.L.slrbs_host_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.slrbs_host_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_bodies(self): return self.L.slrbs_host_num_bodies(self.h)
    @property
    def n_joints(self): return self.L.slrbs_host_num_joints(self.h)
    @property
    def n_contacts(self): return self.L.slrbs_host_num_contacts(self.h)

    def arrays(self):
        """The dense upload arrays of slrbs_upload_bodies / slrbs_upload_joints for this system."""
        n, m = self.n_bodies, self.n_joints
        o = dict(x=np.zeros((n, 3), np.float32), q=np.zeros((n, 4), np.float32), v=np.zeros((n, 3), np.float32),
                 w=np.zeros((n, 3), np.float32), mass=np.zeros(n, np.float32), ibody=np.zeros((n, 9), np.float32),
                 ibodyinv=np.zeros((n, 9), np.float32), fixed=np.zeros(n, np.int32), geom_type=np.zeros(n, np.int32),
                 geom=np.zeros((n, 6), np.float32), jtype=np.zeros(m, np.int32), jb0=np.zeros(m, np.int32),
                 jb1=np.zeros(m, np.int32), r0=np.zeros((m, 3), np.float32), q0=np.zeros((m, 4), np.float32),
                 r1=np.zeros((m, 3), np.float32), q1=np.zeros((m, 4), np.float32), lam=np.zeros((m, 5), np.float32))
        self.L.slrbs_host_export_bodies(self.h, _f(o["x"]), _f(o["q"]), _f(o["v"]), _f(o["w"]), _f(o["mass"]), _f(o["ibody"]),
                                        _f(o["ibodyinv"]), _i(o["fixed"]), _i(o["geom_type"]), _f(o["geom"]))
        self.L.slrbs_host_export_joints(self.h, _i(o["jtype"]), _i(o["jb0"]), _i(o["jb1"]), _f(o["r0"]), _f(o["q0"]),
                                        _f(o["r1"]), _f(o["q1"]), _f(o["lam"]))
        return o

    def set_solver(self, solver_id=0, solver_iter=10, mu=0.8, collisions=True):
        self.L.slrbs_host_set_solver(self.h, int(solver_id), int(solver_iter), float(mu), int(bool(collisions)))

    def set_prestep_force(self, body, f=None, tau=None):
        a = None if f is None else np.ascontiguousarray(np.asarray(f, np.float32))
        b = None if tau is None else np.ascontiguousarray(np.asarray(tau, np.float32))
        self.L.slrbs_host_set_prestep_force(self.h, int(body), _f(a), _f(b))

    def set_state(self, x=None, q=None, v=None, w=None):
        a = [None if t is None else np.ascontiguousarray(np.asarray(t, np.float32)) for t in (x, q, v, w)]
        self.L.slrbs_host_set_state(self.h, *[_f(t) for t in a])

    def step(self, dt=0.01, n=1, device=0):
        if self.L.slrbs_host_step(self.h, float(dt), int(n), int(device)) != 0:
            raise SlrbsError(self.L.slrbs_host_last_error().decode())

    def state(self):
        n = self.n_bodies
        o = dict(x=np.zeros((n, 3), np.float32), q=np.zeros((n, 4), np.float32), v=np.zeros((n, 3), np.float32),
                 w=np.zeros((n, 3), np.float32))
        self.L.slrbs_host_get_state(self.h, _f(o["x"]), _f(o["q"]), _f(o["v"]), _f(o["w"]))
        return o

    def contacts(self):
        m = self.n_contacts
        o = dict(body0=np.zeros(m, np.int32), body1=np.zeros(m, np.int32), p=np.zeros((m, 3), np.float32),
                 n=np.zeros((m, 3), np.float32), phi=np.zeros(m, np.float32), lam=np.zeros((m, 3), np.float32))
        if m:
            self.L.slrbs_host_get_contacts(self.h, _i(o["body0"]), _i(o["body1"]), _f(o["p"]), _f(o["n"]), _f(o["phi"]), _f(o["lam"]))
        return o

    def joint_lambda(self):
        lam = np.zeros((self.n_joints, 5), np.float32)
        if self.n_joints:
            self.L.slrbs_host_get_joint_lambda(self.h, _f(lam))
        return lam

    def save(self): self.L.slrbs_host_save(self.h)
    def restore(self): self.L.slrbs_host_restore(self.h)

    def use_host_solver(self, solver_id=0):
        """Replace solver `solver_id` by a host-side Gauss-Seidel subclass of Solver that only overrides solve(h)."""
        self.L.slrbs_host_use_host_solver(self.h, int(solver_id))


class HostBatch:
    """RigidBodySystemBatch: n_scenes copies of a HostSystem's bodies, joints and current state in one device context."""

    def __init__(self, prototype, n_scenes, device=0, contacts_per_scene=0):
        self.L = load_host_library()
        self.n_scenes, self.n_bodies, self.n_joints = int(n_scenes), prototype.n_bodies, prototype.n_joints
        self.h = self.L.slrbs_host_batch_create(prototype.h, int(n_scenes), int(device), int(contacts_per_scene))
        if not self.h:
            raise SlrbsError(self.L.slrbs_host_last_error().decode())

    def _ck(self, rc):
        if rc != 0:
            raise SlrbsError(self.L.slrbs_host_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.slrbs_host_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_solver(self, solver_id=0, solver_iter=10, mu=0.8, collisions=True):
        self._ck(self.L.slrbs_host_batch_set_solver(self.h, int(solver_id), int(solver_iter), float(mu), int(bool(collisions))))

    def step(self, dt=0.01, n=1): self._ck(self.L.slrbs_host_batch_step(self.h, float(dt), int(n)))
    def sync(self): self._ck(self.L.slrbs_host_batch_sync(self.h))
    def save(self): self._ck(self.L.slrbs_host_batch_save(self.h))
    def reset(self, scene0=0, n_scenes=None): self._ck(self.L.slrbs_host_batch_reset(self.h, int(scene0), int(n_scenes or self.n_scenes - scene0)))

    def state(self, scene0=0, n_scenes=None):
        ns, n = int(n_scenes or self.n_scenes - scene0), self.n_bodies
        o = dict(x=np.zeros((ns, n, 3), np.float32), q=np.zeros((ns, n, 4), np.float32), v=np.zeros((ns, n, 3), np.float32),
                 w=np.zeros((ns, n, 3), np.float32))
        self._ck(self.L.slrbs_host_batch_get_state(self.h, int(scene0), ns, _f(o["x"]), _f(o["q"]), _f(o["v"]), _f(o["w"])))
        return o

    def set_state(self, x, q, v, w, scene0=0):
        a = [np.ascontiguousarray(np.asarray(t, np.float32)) for t in (x, q, v, w)]
        self._ck(self.L.slrbs_host_batch_set_state(self.h, int(scene0), a[0].shape[0], *[_f(t) for t in a]))

    def set_forces(self, f=None, tau=None, scene0=0, n_scenes=1, absolute=False):
        a = [None if t is None else np.ascontiguousarray(np.asarray(t, np.float32)) for t in (f, tau)]
        self._ck(self.L.slrbs_host_batch_set_forces(self.h, int(scene0), int(n_scenes), _f(a[0]), _f(a[1]), int(bool(absolute))))

    def contact_counts(self):
        c = np.zeros(self.n_scenes, np.int32)
        self._ck(self.L.slrbs_host_batch_contact_counts(self.h, 0, self.n_scenes, _i(c)))
        return c

    def read_scene(self, scene, into):
        self._ck(self.L.slrbs_host_batch_read_scene(self.h, int(scene), into.h))


def scene_arrays(scene, a=0, b=0, c=0):
    """Upload arrays of one of the Scenarios.h scenes (built by the C++ mirror, no GPU needed)."""
    s = HostSystem(scene, a, b, c)
    try:
        return s.arrays()
    finally:
        s.close()
