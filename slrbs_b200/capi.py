"""ctypes binding of include/slrbs_b200.h (libslrbs_b200.so).

Fails loudly when the CUDA library is missing or when no sm_100 device is usable; there is
no CPU path in the product.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SPHERE, BOX, PLANE, CYLINDER = 0, 1, 2, 3
SPHERICAL, HINGE = 1, 2

# every symbol include/slrbs_b200.h declares
SYMBOLS = [
    "slrbs_last_error", "slrbs_version", "slrbs_describe", "slrbs_create", "slrbs_destroy", "slrbs_set_params",
    "slrbs_upload_bodies", "slrbs_upload_joints", "slrbs_upload_state", "slrbs_download_state", "slrbs_download_state_async", "slrbs_upload_state_async",
    "slrbs_set_external_forces", "slrbs_step", "slrbs_sync", "slrbs_stage_prepare", "slrbs_stage_detect",
    "slrbs_stage_jacobians", "slrbs_stage_solve", "slrbs_stage_integrate", "slrbs_download_contact_counts",
    "slrbs_download_contacts", "slrbs_download_contact_rows", "slrbs_download_joint_rows",
    "slrbs_download_joint_lambda", "slrbs_upload_joint_lambda", "slrbs_upload_contact_lambda", "slrbs_download_contact_schedule", "slrbs_download_body_dyn",
    "slrbs_save_state", "slrbs_restore_state", "slrbs_profile_enable", "slrbs_profile_read",
    "slrbs_timer_start", "slrbs_timer_stop", "slrbs_halo_set_local", "slrbs_halo_set_remote", "slrbs_halo_sync_local",
    "slrbs_halo_pack", "slrbs_halo_unpack", "slrbs_stream", "slrbs_device", "slrbs_pile_init", "slrbs_pile_export", "slrbs_pile_rebuild",
    "slrbs_pile_requests", "slrbs_pile_set_send", "slrbs_pile_scene_keys", "slrbs_set_extensions",
    "slrbs_sdf_build", "slrbs_sdf_add_shape", "slrbs_debug_div_rn",
]


class SlrbsError(RuntimeError):
    pass


def lib_path():
    # SLRBS_LIB lets a developer point the binding at an experimental build of the same ABI
    return os.environ.get("SLRBS_LIB") or os.path.join(_HERE, "libslrbs_b200.so")


def load_library():
    """dlopen the C ABI library; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise SlrbsError(f"{p} is missing: build it with `make -C slrbs_b200/csrc` "
                         "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(p)
    fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_void_p
    L.slrbs_last_error.restype = C.c_char_p
    L.slrbs_version.restype = C.c_char_p
    L.slrbs_describe.argtypes = [vp, C.c_char_p, C.c_int]
    L.slrbs_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.slrbs_destroy.argtypes = [vp]
    L.slrbs_destroy.restype = None
    L.slrbs_set_params.argtypes = [vp, C.c_float, C.c_int, C.c_int, C.c_int]
    L.slrbs_set_extensions.argtypes = [vp, C.c_int]
    L.slrbs_sdf_build.argtypes = [C.c_int, C.c_int, fp, C.c_int, ip, C.c_int, C.c_int, C.c_int, fp, C.c_float, fp]
    L.slrbs_sdf_add_shape.argtypes = [vp, C.c_int, C.c_int, C.c_int, fp, C.c_float, fp, C.c_int, fp]
    L.slrbs_upload_bodies.argtypes = [vp, C.c_int, C.c_int, C.c_int, ip, fp, fp, fp, fp, fp, fp, fp, ip, ip, fp]
    L.slrbs_upload_joints.argtypes = [vp, C.c_int, C.c_int, C.c_int, ip, ip, ip, ip, fp, fp, fp, fp]
    L.slrbs_upload_state.argtypes = [vp, C.c_int, C.c_int, C.c_int, fp, fp, fp, fp]
    L.slrbs_download_state_async.argtypes = [vp, C.c_int, C.c_int, C.c_int, fp, fp, fp, fp]
    L.slrbs_upload_state_async.argtypes = [vp, C.c_int, C.c_int, C.c_int, fp, fp, fp, fp]
    L.slrbs_download_state.argtypes = [vp, C.c_int, C.c_int, C.c_int, fp, fp, fp, fp]
    L.slrbs_set_external_forces.argtypes = [vp, C.c_int, C.c_int, C.c_int, fp, fp, C.c_int]
    L.slrbs_step.argtypes = [vp, C.c_float, C.c_int]
    L.slrbs_sync.argtypes = [vp]
    L.slrbs_stage_prepare.argtypes = [vp]
    L.slrbs_stage_detect.argtypes = [vp]
    L.slrbs_stage_jacobians.argtypes = [vp, C.c_float]
    L.slrbs_stage_solve.argtypes = [vp, C.c_float]
    L.slrbs_stage_integrate.argtypes = [vp, C.c_float]
    L.slrbs_download_contact_counts.argtypes = [vp, C.c_int, C.c_int, ip]
    L.slrbs_download_contacts.argtypes = [vp, C.c_int, C.c_int, ip, ip, fp, fp, fp, fp, fp, fp]
    L.slrbs_download_contact_rows.argtypes = [vp, C.c_int, C.c_int, fp, fp, fp, fp, fp, fp]
    L.slrbs_download_joint_rows.argtypes = [vp, C.c_int, C.c_int, fp, fp, fp, fp, fp]
    L.slrbs_download_joint_lambda.argtypes = [vp, C.c_int, C.c_int, C.c_int, fp]
    L.slrbs_upload_joint_lambda.argtypes = [vp, C.c_int, C.c_int, C.c_int, fp]
    L.slrbs_upload_contact_lambda.argtypes = [vp, C.c_int, C.c_int, fp]
    L.slrbs_download_contact_schedule.argtypes = [vp, C.c_int, C.c_int, ip, C.c_int, ip, ip]
    L.slrbs_download_body_dyn.argtypes = [vp, C.c_int, C.c_int, fp, fp, fp, fp, fp, fp]
    L.slrbs_save_state.argtypes = [vp]
    L.slrbs_restore_state.argtypes = [vp, C.c_int, C.c_int]
    L.slrbs_profile_enable.argtypes = [vp, C.c_int]
    L.slrbs_halo_set_local.argtypes = [vp, C.c_int, ip, ip]
    L.slrbs_halo_set_remote.argtypes = [vp, C.c_int, ip, C.c_int, ip]
    L.slrbs_halo_sync_local.argtypes = [vp]
    L.slrbs_halo_pack.argtypes = [vp, C.c_void_p]
    L.slrbs_stream.argtypes = [vp]
    L.slrbs_stream.restype = C.c_void_p
    L.slrbs_device.argtypes = [vp]
    L.slrbs_debug_div_rn.argtypes = [fp, fp, fp, C.c_int]
    L.slrbs_halo_unpack.argtypes = [vp, C.c_void_p]
    L.slrbs_pile_init.argtypes = [vp, C.c_int, fp, fp, fp, C.c_float, C.c_float, ip, C.c_int, C.c_int, C.c_int, fp]
    L.slrbs_pile_export.argtypes = [vp, C.c_void_p, C.c_void_p]
    L.slrbs_pile_rebuild.argtypes = [vp, C.c_void_p, ip]
    L.slrbs_pile_requests.argtypes = [vp, C.c_void_p]
    L.slrbs_pile_set_send.argtypes = [vp, C.c_void_p, C.c_int]
    L.slrbs_pile_scene_keys.argtypes = [vp, C.c_int, C.c_int, ip]
    L.slrbs_timer_start.argtypes = [vp]
    L.slrbs_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.slrbs_profile_read.argtypes = [vp, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                     C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    _LIB = L
    return L


def _f(a):
    return None if a is None else C.cast(a.ctypes.data, C.POINTER(C.c_float))


def _i(a):
    return None if a is None else C.cast(a.ctypes.data, C.POINTER(C.c_int32))


def _arr(a, dtype, shape):
    if a is None:
        return None
    a = np.ascontiguousarray(np.asarray(a, dtype=dtype))
    if a.shape != tuple(shape):
        a = np.ascontiguousarray(np.broadcast_to(a.reshape((-1,) + tuple(shape[1:])) if a.ndim == len(shape) - 1 else a, shape))
    return a


class Context:
    """A batch of n_scenes independent RigidBodySystems resident on one B200."""

    def __init__(self, n_scenes, max_bodies, max_joints=0, max_contacts=0, device=0):
        self.L = load_library()
        self.h = C.c_void_p()
        self.n_scenes, self.max_bodies, self.max_joints, self.max_contacts = n_scenes, max_bodies, max_joints, max_contacts
        self.n_bodies = max_bodies
        self.n_joints = 0
        self._ck(self.L.slrbs_create(C.byref(self.h), device, n_scenes, max_bodies, max_joints, max_contacts))

    def _ck(self, rc):
        if rc < 0:
            raise SlrbsError(f"slrbs error {rc}: {self.L.slrbs_last_error().decode()}")
        return rc

    def set_extensions(self, flags=1):
        """Extension contact pairs without a reference counterpart (sphere-plane, box-plane, box-box)."""
        self._ck(self.L.slrbs_set_extensions(self.h, int(flags)))

    def add_sdf_shape(self, dims, origin, cell, grid, verts):
        """Extension: a mesh as signed-distance grid + sample vertices; returns the shape id for geom_type 4 bodies."""
        g = np.ascontiguousarray(grid, np.float32).reshape(-1); v = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
        o = np.ascontiguousarray(origin, np.float32)
        assert g.size == dims[0] * dims[1] * dims[2]
        return self._ck(self.L.slrbs_sdf_add_shape(self.h, int(dims[0]), int(dims[1]), int(dims[2]), _f(o), float(cell), _f(g), len(v), _f(v)))

    def describe(self):
        buf = C.create_string_buffer(256)
        self._ck(self.L.slrbs_describe(self.h, buf, 256))
        return buf.value.decode()

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.slrbs_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setup ---------------------------------------------------------------------
    def set_params(self, mu=0.8, solver_id=0, solver_iters=10, collisions=True):
        self._ck(self.L.slrbs_set_params(self.h, float(mu), int(solver_id), int(solver_iters), int(bool(collisions))))

    def upload_bodies(self, x, q, v, w, mass, ibody, ibodyinv, fixed, geom_type, geom, scene0=0, counts=None):
        """Arrays are [n_scenes, n_bodies, K] (or [n_bodies, K], replicated to every scene)."""
        x = np.asarray(x, np.float32)
        ns = self.n_scenes - scene0 if x.ndim == 2 else x.shape[0]
        n = x.shape[-2]

        def rep(a, dt, k):
            a = np.asarray(a, dt)
            tgt = (ns, n) + ((k,) if k else ())
            if a.shape != tgt:
                a = np.broadcast_to(a.reshape((1, n) + ((k,) if k else ())), tgt)
            return np.ascontiguousarray(a)

        a = [rep(x, np.float32, 3), rep(q, np.float32, 4), rep(v, np.float32, 3), rep(w, np.float32, 3),
             rep(mass, np.float32, 0), rep(ibody, np.float32, 9), rep(ibodyinv, np.float32, 9)]
        fx, gt, g = rep(fixed, np.int32, 0), rep(geom_type, np.int32, 0), rep(geom, np.float32, 6)
        cnt = None if counts is None else np.ascontiguousarray(np.asarray(counts, np.int32))
        self._ck(self.L.slrbs_upload_bodies(self.h, scene0, ns, n, _i(cnt), *[_f(t) for t in a], _i(fx), _i(gt), _f(g)))
        self.n_bodies = n

    def upload_joints(self, jtype, body0, body1, r0, q0, r1, q1, scene0=0, counts=None):
        jtype = np.asarray(jtype, np.int32)
        ns = self.n_scenes - scene0 if jtype.ndim == 1 else jtype.shape[0]
        n = jtype.shape[-1]
        if n == 0:
            self._ck(self.L.slrbs_upload_joints(self.h, scene0, ns, 0, None, None, None, None, None, None, None, None))
            self.n_joints = 0
            return

        def rep(a, dt, k):
            a = np.asarray(a, dt)
            tgt = (ns, n) + ((k,) if k else ())
            if a.shape != tgt:
                a = np.broadcast_to(a.reshape((1, n) + ((k,) if k else ())), tgt)
            return np.ascontiguousarray(a)

        ty, b0, b1 = rep(jtype, np.int32, 0), rep(body0, np.int32, 0), rep(body1, np.int32, 0)
        a = [rep(r0, np.float32, 3), rep(q0, np.float32, 4), rep(r1, np.float32, 3), rep(q1, np.float32, 4)]
        cnt = None if counts is None else np.ascontiguousarray(np.asarray(counts, np.int32))
        self._ck(self.L.slrbs_upload_joints(self.h, scene0, ns, n, _i(cnt), _i(ty), _i(b0), _i(b1), *[_f(t) for t in a]))
        self.n_joints = n

    def upload_state(self, x=None, q=None, v=None, w=None, scene0=0, n_scenes=None, n_bodies=None):
        ns = n_scenes or (self.n_scenes - scene0)
        n = n_bodies or self.n_bodies
        a = [None if t is None else np.ascontiguousarray(np.asarray(t, np.float32).reshape(ns, n, k))
             for t, k in ((x, 3), (q, 4), (v, 3), (w, 3))]
        self._ck(self.L.slrbs_upload_state(self.h, scene0, ns, n, *[_f(t) for t in a]))

    def upload_state_raw(self, ns, n, px, pq, pv, pw, scene0=0):
        """Raw host pointers (ints), e.g. pinned torch tensors' data_ptr()."""
        cast = lambda p: None if not p else C.cast(p, C.POINTER(C.c_float))
        self._ck(self.L.slrbs_upload_state(self.h, scene0, ns, n, cast(px), cast(pq), cast(pv), cast(pw)))

    def download_state_raw(self, ns, n, px, pq, pv, pw, scene0=0):
        cast = lambda p: None if not p else C.cast(p, C.POINTER(C.c_float))
        self._ck(self.L.slrbs_download_state(self.h, scene0, ns, n, cast(px), cast(pq), cast(pv), cast(pw)))

    def upload_state_raw_async(self, ns, n, px, pq, pv, pw, scene0=0):
        """No wait: the (pinned) host buffers must not change before the next sync()."""
        cast = lambda p: None if not p else C.cast(p, C.POINTER(C.c_float))
        self._ck(self.L.slrbs_upload_state_async(self.h, scene0, ns, n, cast(px), cast(pq), cast(pv), cast(pw)))

    def download_state_raw_async(self, ns, n, px, pq, pv, pw, scene0=0):
        """No wait: the (pinned) host buffers are valid after the next sync()."""
        cast = lambda p: None if not p else C.cast(p, C.POINTER(C.c_float))
        self._ck(self.L.slrbs_download_state_async(self.h, scene0, ns, n, cast(px), cast(pq), cast(pv), cast(pw)))

    def set_external_forces_raw(self, ns, n, pf, pt, scene0=0, absolute=False):
        """Raw host pointers (ints) to [ns][n][3] force / torque arrays, e.g. pinned torch tensors' data_ptr()."""
        cast = lambda p: None if not p else C.cast(p, C.POINTER(C.c_float))
        self._ck(self.L.slrbs_set_external_forces(self.h, scene0, ns, n, cast(pf), cast(pt), int(bool(absolute))))

    def download_state(self, scene0=0, n_scenes=None, n_bodies=None):
        ns = n_scenes or (self.n_scenes - scene0)
        n = n_bodies or self.n_bodies
        x = np.zeros((ns, n, 3), np.float32); q = np.zeros((ns, n, 4), np.float32)
        v = np.zeros((ns, n, 3), np.float32); w = np.zeros((ns, n, 3), np.float32)
        self._ck(self.L.slrbs_download_state(self.h, scene0, ns, n, _f(x), _f(q), _f(v), _f(w)))
        return dict(x=x, q=q, v=v, w=w)

    def set_external_forces(self, f=None, tau=None, scene0=0, n_scenes=None, n_bodies=None, absolute=False):
        ns = n_scenes or (self.n_scenes - scene0)
        n = n_bodies or self.n_bodies
        a = [None if t is None else np.ascontiguousarray(np.broadcast_to(np.asarray(t, np.float32).reshape(-1, n, 3), (ns, n, 3)))
             for t in (f, tau)]
        self._ck(self.L.slrbs_set_external_forces(self.h, scene0, ns, n, _f(a[0]), _f(a[1]), int(bool(absolute))))

    # -- stepping ------------------------------------------------------------------
    def step(self, dt=0.01, n=1):
        self._ck(self.L.slrbs_step(self.h, float(dt), int(n)))

    def sync(self):
        self._ck(self.L.slrbs_sync(self.h))

    def stage_prepare(self): self._ck(self.L.slrbs_stage_prepare(self.h))
    def stage_detect(self): self._ck(self.L.slrbs_stage_detect(self.h))
    def stage_jacobians(self, dt): self._ck(self.L.slrbs_stage_jacobians(self.h, float(dt)))
    def stage_solve(self, dt): self._ck(self.L.slrbs_stage_solve(self.h, float(dt)))
    def stage_integrate(self, dt): self._ck(self.L.slrbs_stage_integrate(self.h, float(dt)))
    def save_state(self): self._ck(self.L.slrbs_save_state(self.h))

    def restore_state(self, scene0=0, n_scenes=None):
        self._ck(self.L.slrbs_restore_state(self.h, scene0, n_scenes or (self.n_scenes - scene0)))

    # -- read back -----------------------------------------------------------------
    def contact_counts(self, scene0=0, n_scenes=None):
        ns = n_scenes or (self.n_scenes - scene0)
        c = np.zeros(ns, np.int32)
        self._ck(self.L.slrbs_download_contact_counts(self.h, scene0, ns, _i(c)))
        return c

    def contacts(self, scene=0):
        m = max(int(self.contact_counts(scene, 1)[0]), 1)
        b0 = np.zeros(m, np.int32); b1 = np.zeros(m, np.int32)
        o = {k: np.zeros((m, 3), np.float32) for k in ("p", "n", "t", "b", "lam")}
        phi = np.zeros(m, np.float32)
        cnt = self._ck(self.L.slrbs_download_contacts(self.h, scene, m, _i(b0), _i(b1), _f(o["p"]), _f(o["n"]),
                                                      _f(o["t"]), _f(o["b"]), _f(phi), _f(o["lam"])))
        out = dict(body0=b0[:cnt], body1=b1[:cnt], phi=phi[:cnt])
        out.update({k: a[:cnt] for k, a in o.items()})
        return out

    def contact_rows(self, scene=0):
        m = max(int(self.contact_counts(scene, 1)[0]), 1)
        o = {k: np.zeros((m, 3, 6), np.float32) for k in ("J0", "J1", "J0Minv", "J1Minv")}
        A = np.zeros((m, 3, 3), np.float32); rhs = np.zeros((m, 3), np.float32)
        cnt = self._ck(self.L.slrbs_download_contact_rows(self.h, scene, m, _f(o["J0"]), _f(o["J1"]), _f(o["J0Minv"]),
                                                          _f(o["J1Minv"]), _f(A), _f(rhs)))
        out = {k: a[:cnt] for k, a in o.items()}
        out["A"] = A[:cnt]; out["rhs"] = rhs[:cnt]
        return out

    def joint_rows(self, scene=0):
        m = max(self.n_joints, 1)
        o = {k: np.zeros((m, 5, 6), np.float32) for k in ("J0", "J1", "J0Minv", "J1Minv")}
        rhs = np.zeros((m, 5), np.float32)
        cnt = self._ck(self.L.slrbs_download_joint_rows(self.h, scene, m, _f(o["J0"]), _f(o["J1"]), _f(o["J0Minv"]),
                                                        _f(o["J1Minv"]), _f(rhs)))
        out = {k: a[:cnt] for k, a in o.items()}
        out["rhs"] = rhs[:cnt]
        return out

    def joint_lambda(self, scene0=0, n_scenes=None):
        ns = n_scenes or (self.n_scenes - scene0)
        lam = np.zeros((ns, self.n_joints, 5), np.float32)
        if self.n_joints:
            self._ck(self.L.slrbs_download_joint_lambda(self.h, scene0, ns, self.n_joints, _f(lam)))
        return lam

    def set_joint_lambda(self, lam, scene0=0):
        lam = np.ascontiguousarray(np.asarray(lam, np.float32).reshape(-1, self.n_joints, 5))
        self._ck(self.L.slrbs_upload_joint_lambda(self.h, scene0, lam.shape[0], self.n_joints, _f(lam)))

    def contact_schedule(self, scene=0):
        """(position of every contact in the solver's row stream, end of every dependency level / colour in it)."""
        n = int(self.contact_counts(scene, 1)[0])
        pos = np.zeros(max(n, 1), np.int32); ends = np.zeros(max(n, 64), np.int32); nl = np.zeros(1, np.int32)
        self._ck(self.L.slrbs_download_contact_schedule(self.h, scene, n, _i(pos), len(ends), _i(ends), _i(nl)))
        return pos[:n], ends[:int(nl[0])]

    def set_contact_lambda(self, lam, scene=0):
        """Contact::lambda of one scene in detection order (what a host-side Solver subclass leaves behind)."""
        lam = np.ascontiguousarray(np.asarray(lam, np.float32).reshape(-1, 3))
        self._ck(self.L.slrbs_upload_contact_lambda(self.h, scene, lam.shape[0], _f(lam)))

    def body_dyn(self, scene=0):
        n = self.n_bodies
        o = {k: np.zeros((n, 3), np.float32) for k in ("f", "tau", "fc", "tauc")}
        o["I"] = np.zeros((n, 9), np.float32); o["Iinv"] = np.zeros((n, 9), np.float32)
        self._ck(self.L.slrbs_download_body_dyn(self.h, scene, n, _f(o["f"]), _f(o["tau"]), _f(o["fc"]), _f(o["tauc"]),
                                                _f(o["I"]), _f(o["Iinv"])))
        return o

    # -- halos of a scene cut into blocks -------------------------------------------
    def halo_set_local(self, src_flat, dst_flat):
        a = np.ascontiguousarray(np.asarray(src_flat, np.int32)); b = np.ascontiguousarray(np.asarray(dst_flat, np.int32))
        self._ck(self.L.slrbs_halo_set_local(self.h, a.size, _i(a), _i(b)))

    def halo_set_remote(self, send_flat, recv_flat):
        a = np.ascontiguousarray(np.asarray(send_flat, np.int32)); b = np.ascontiguousarray(np.asarray(recv_flat, np.int32))
        self._ck(self.L.slrbs_halo_set_remote(self.h, a.size, _i(a), b.size, _i(b)))

    def halo_sync_local(self): self._ck(self.L.slrbs_halo_sync_local(self.h))
    def halo_pack(self, device_ptr): self._ck(self.L.slrbs_halo_pack(self.h, C.c_void_p(device_ptr)))
    def halo_unpack(self, device_ptr): self._ck(self.L.slrbs_halo_unpack(self.h, C.c_void_p(device_ptr)))

    # -- one large pile re-partitioned on the device --------------------------------
    def pile_init(self, radius, mass, origin, block, halo_w, ncell, cx0, cx1, fixed10):
        r = np.ascontiguousarray(np.asarray(radius, np.float32)); m = np.ascontiguousarray(np.asarray(mass, np.float32))
        o = np.ascontiguousarray(np.asarray(origin, np.float32)); nc = np.ascontiguousarray(np.asarray(ncell, np.int32))
        f = np.ascontiguousarray(np.asarray(fixed10, np.float32).reshape(-1, 10))
        self._ck(self.L.slrbs_pile_init(self.h, r.size, _f(r), _f(m), _f(o), float(block), float(halo_w), _i(nc), int(cx0), int(cx1),
                                        f.shape[0], _f(f) if f.shape[0] else None))

    def pile_export(self, table_ptr, vmax2_ptr):
        self._ck(self.L.slrbs_pile_export(self.h, C.c_void_p(table_ptr), C.c_void_p(vmax2_ptr)))

    def pile_rebuild(self, table_ptr):
        """Returns info8; raises SlrbsError on capacity overflow (info8 is then in self.pile_info)."""
        info = np.zeros(8, np.int32)
        self.pile_info = info
        self._ck(self.L.slrbs_pile_rebuild(self.h, C.c_void_p(table_ptr), _i(info)))
        return info

    def pile_requests(self, gids_ptr): self._ck(self.L.slrbs_pile_requests(self.h, C.c_void_p(gids_ptr)))
    def pile_set_send(self, gids_ptr, n): self._ck(self.L.slrbs_pile_set_send(self.h, C.c_void_p(gids_ptr), int(n)))

    def pile_scene_keys(self, scene):
        out = np.zeros(self.max_bodies, np.int32)
        n = self.L.slrbs_pile_scene_keys(self.h, int(scene), out.size, _i(out))
        if n < 0:
            self._ck(n)
        return out[:n]

    # -- measurement ---------------------------------------------------------------
    def profile_enable(self, on=True):
        self._ck(self.L.slrbs_profile_enable(self.h, int(bool(on))))

    def timer_start(self):
        self._ck(self.L.slrbs_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self.L.slrbs_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def profile_read(self, reset=True):
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        ms = C.c_double()
        self._ck(self.L.slrbs_profile_read(self.h, int(bool(reset)), C.byref(a), C.byref(ms), C.byref(b), C.byref(c), C.byref(d)))
        return dict(solve_launches=a.value, solve_ms=ms.value, contact_visits=b.value, joint_visits=c.value,
                    total_launches=d.value)


def sdf_build(verts, tris, dims, origin, cell, device=0):
    """Extension: signed-distance grid of a closed triangle mesh, computed on the device (slrbs_sdf_build)."""
    L = load_library()
    v = np.ascontiguousarray(verts, np.float32).reshape(-1, 3); t = np.ascontiguousarray(tris, np.int32).reshape(-1, 3)
    o = np.ascontiguousarray(origin, np.float32)
    out = np.zeros(int(dims[0]) * int(dims[1]) * int(dims[2]), np.float32)
    rc = L.slrbs_sdf_build(int(device), len(v), _f(v), len(t), _i(t), int(dims[0]), int(dims[1]), int(dims[2]), _f(o), float(cell), _f(out))
    if rc < 0:
        raise SlrbsError(L.slrbs_last_error().decode())
    return out


def debug_div_rn(a, b):
    """Test hook (slrbs_debug_div_rn): a / b evaluated on the device by the row solve's call-free division."""
    L = load_library()
    a = np.ascontiguousarray(a, np.float32).ravel(); b = np.ascontiguousarray(b, np.float32).ravel()
    if a.shape != b.shape:
        raise ValueError("a and b must have the same number of elements")
    out = np.zeros_like(a)
    rc = L.slrbs_debug_div_rn(_f(a), _f(b), _f(out), int(a.size))
    if rc < 0:
        raise SlrbsError(L.slrbs_last_error().decode())
    return out
