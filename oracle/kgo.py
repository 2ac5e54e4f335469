"""ctypes binding of the fp64 oracle (oracle/libkgo.so).  TEST INFRASTRUCTURE ONLY:
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never from kinovagrasping_b200/.  See oracle/kgo.h for the parity-unpinned statement."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NQ, NV, NU, NBODY, NGEOM, NSITE, NSENSOR, MAXCON, NOBS = 16, 15, 9, 9, 9, 17, 26, 64, 82
MAXEFC = 3 + 6 + 4 * MAXCON
BLOB = os.path.join(HERE, "..", "kinovagrasping_b200", "assets", "model_card.bin")
SHAPES = ["CubeS", "CubeB", "CylinderS", "CylinderB", "Cube45S", "Cube45B", "Cone1S", "Cone1B",
          "Cone2S", "Cone2B", "Vase1S", "Vase1B", "Vase2S", "Vase2B"]

d = C.c_double


class Contact(C.Structure):
    _fields_ = [("geom1", C.c_int), ("geom2", C.c_int), ("dist", d), ("pos", d * 3), ("frame", d * 9),
                ("friction", d), ("includemargin", d)]


class Data(C.Structure):
    _fields_ = [("xpos", d * (NBODY * 3)), ("xmat", d * (NBODY * 9)), ("xipos", d * (NBODY * 3)), ("ximat", d * (NBODY * 9)),
                ("geom_xpos", d * (NGEOM * 3)), ("geom_xmat", d * (NGEOM * 9)),
                ("site_xpos", d * (NSITE * 3)), ("site_zaxis", d * (NSITE * 3)),
                ("M", d * (NV * NV)),
                ("qfrc_bias", d * NV), ("qfrc_passive", d * NV), ("qfrc_actuator", d * NV), ("qfrc_smooth", d * NV),
                ("qacc_smooth", d * NV), ("qacc", d * NV), ("qfrc_constraint", d * NV),
                ("ncon", C.c_int), ("con", Contact * MAXCON), ("nefc", C.c_int),
                ("efc_J", d * (MAXEFC * NV)), ("efc_pos", d * MAXEFC), ("efc_margin", d * MAXEFC),
                ("efc_D", d * MAXEFC), ("efc_R", d * MAXEFC), ("efc_aref", d * MAXEFC), ("efc_force", d * MAXEFC),
                ("efc_type", C.c_int * MAXEFC), ("sensordata", d * NSENSOR),
                ("solver_iter", C.c_int), ("n_mpr", C.c_int), ("n_broad", C.c_int), ("solver_cost", d)]

    def arr(self, name, shape=None):
        a = np.ctypeslib.as_array(getattr(self, name)).copy()
        return a.reshape(shape) if shape else a


class State(C.Structure):
    _fields_ = [("qpos", d * NQ), ("qvel", d * NV), ("qacc_warmstart", d * NV), ("hand_quat", d * 4), ("shape", C.c_int)]


class Env(C.Structure):
    _fields_ = [("st", State), ("d", Data), ("elapsed", C.c_int), ("max_steps", C.c_int),
                ("init_obj_x", d), ("init_dot", d), ("prev_fx", d * 3),
                ("have_prev", C.c_int), ("have_prev2", C.c_int), ("num_good_grasps", C.c_int), ("total_steps", C.c_int),
                ("obs0_fx", d * 3)]


def build(force=False):
    so = os.path.join(HERE, "libkgo.so")
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", HERE, "libkgo.so"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.kgo_model_load.restype = C.c_void_p
        L.kgo_model_load.argtypes = [C.c_void_p, C.c_size_t]
        L.kgo_state_init.argtypes = [C.c_void_p, C.POINTER(State), C.c_int, C.POINTER(d), C.POINTER(d)]
        L.kgo_forward.argtypes = [C.c_void_p, C.POINTER(State), C.POINTER(d), C.c_int, C.POINTER(Data)]
        L.kgo_forward_contacts.argtypes = [C.c_void_p, C.POINTER(State), C.POINTER(d), C.c_int, C.c_int, C.POINTER(d), C.POINTER(Data)]
        L.kgo_step_contacts.argtypes = [C.c_void_p, C.POINTER(State), C.POINTER(d), C.c_int, C.c_int, C.POINTER(d), C.POINTER(Data)]
        L.kgo_step.argtypes = [C.c_void_p, C.POINTER(State), C.POINTER(d), C.c_int, C.POINTER(Data)]
        L.kgo_env_reset.argtypes = [C.c_void_p, C.POINTER(Env), C.c_int, C.POINTER(d), C.POINTER(d), C.POINTER(d)]
        L.kgo_env_step.argtypes = [C.c_void_p, C.POINTER(Env), C.POINTER(d), C.POINTER(d), C.POINTER(d), C.POINTER(d)]
        L.kgo_env_step.restype = C.c_int
        L.kgo_env_reset_pose.argtypes = [C.c_void_p, C.POINTER(Env), C.c_int, C.POINTER(d), C.POINTER(d), C.POINTER(d), C.POINTER(d)]
        L.kgo_env_step6.argtypes = [C.c_void_p, C.POINTER(Env), C.POINTER(d), C.POINTER(d), C.POINTER(d), C.POINTER(d)]
        L.kgo_env_step6.restype = C.c_int
        L.kgo_ctrl_from_action6.argtypes = [C.POINTER(Data), C.POINTER(d), C.POINTER(d)]
        L.kgo_hand_location.argtypes = [C.POINTER(Data), C.c_int, C.POINTER(d)]
        L.kgo_get_obs.argtypes = [C.c_void_p, C.POINTER(State), C.POINTER(Data), C.c_int, C.POINTER(d)]
        L.kgo_ctrl_from_action.argtypes = [C.POINTER(Data), C.POINTER(d), C.POINTER(d)]
        L.kgo_expert_action.argtypes = [C.POINTER(Env), C.POINTER(d), C.c_int, C.POINTER(d), C.POINTER(C.c_int)]
        L.kgo_mpr.argtypes = [C.c_void_p, C.c_int, C.POINTER(d), C.POINTER(d), C.c_int, C.POINTER(d), C.POINTER(d), d,
                              C.POINTER(d), C.POINTER(d), C.POINTER(d)]
        L.kgo_mpr.restype = C.c_int
        L.kgo_ray_mesh.argtypes = [C.c_void_p, C.c_int, C.POINTER(d), C.POINTER(d), C.POINTER(d), C.POINTER(d), C.POINTER(d), C.c_int]
        L.kgo_ray_mesh.restype = d
        L.kgo_model_section.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.kgo_model_section.restype = C.POINTER(d)
        assert L.kgo_sizeof_data() == C.sizeof(Data), (L.kgo_sizeof_data(), C.sizeof(Data))
        assert L.kgo_sizeof_state() == C.sizeof(State)
        assert L.kgo_sizeof_env() == C.sizeof(Env)
        _lib = L
    return _lib


def _p(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(d))


class Model:
    def __init__(self, blob_path=BLOB):
        raw = open(blob_path, "rb").read()
        self._raw = raw
        self.h = lib().kgo_model_load(raw, len(raw))
        assert self.h, "bad model blob"

    def section(self, name):
        n = C.c_int()
        p = lib().kgo_model_section(self.h, name.encode(), C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def state(self, shape, hand_euler, obj_pos):
        s = State()
        _, e = _p(hand_euler); _, o = _p(obj_pos)
        lib().kgo_state_init(self.h, C.byref(s), shape, e, o)
        return s

    def forward(self, s, ctrl=None, rays=False):
        dat = Data()
        c, cp = _p(np.zeros(NU) if ctrl is None else ctrl)
        lib().kgo_forward(self.h, C.byref(s), cp, 1 if rays else 0, C.byref(dat))
        return dat

    def forward_contacts(self, s, ctrl, contacts):
        """mj_forward with a GIVEN contact list: rows of (geom1, geom2, dist, pos[3], normal[3]) -- kgo_forward_contacts."""
        dat = Data()
        c, cp = _p(np.zeros(NU) if ctrl is None else ctrl)
        con = np.ascontiguousarray(contacts, dtype=np.float64).reshape(-1, 9)
        g, gp = _p(con.ravel() if len(con) else np.zeros(9))
        lib().kgo_forward_contacts(self.h, C.byref(s), cp, 0, len(con), gp, C.byref(dat))
        return dat

    def step_contacts(self, s, ctrl, contacts):
        """one mj_step (s is advanced in place) with a GIVEN contact list -- kgo_step_contacts."""
        dat = Data()
        c, cp = _p(np.zeros(NU) if ctrl is None else ctrl)
        con = np.ascontiguousarray(contacts, dtype=np.float64).reshape(-1, 9)
        g, gp = _p(con.ravel() if len(con) else np.zeros(9))
        lib().kgo_step_contacts(self.h, C.byref(s), cp, 0, len(con), gp, C.byref(dat))
        return dat

    def step(self, s, ctrl=None, rays=False):
        dat = Data()
        c, cp = _p(np.zeros(NU) if ctrl is None else ctrl)
        lib().kgo_step(self.h, C.byref(s), cp, 1 if rays else 0, C.byref(dat))
        return dat

    def get_obs(self, s, dat, global_rep=False):
        out = np.zeros(82)
        lib().kgo_get_obs(self.h, C.byref(s), C.byref(dat), 1 if global_rep else 0, out.ctypes.data_as(C.POINTER(d)))
        return out[:74] if global_rep else out

    def env_reset(self, shape, hand_euler, obj_pos, hand_pos=None):
        e = Env()
        obs = np.zeros(NOBS)
        _, eu = _p(hand_euler); _, o = _p(obj_pos)
        if hand_pos is None:
            lib().kgo_env_reset(self.h, C.byref(e), shape, eu, o, obs.ctypes.data_as(C.POINTER(d)))
        else:
            _, hp = _p(hand_pos)
            lib().kgo_env_reset_pose(self.h, C.byref(e), shape, eu, o, hp, obs.ctypes.data_as(C.POINTER(d)))
        return e, obs

    def hand_location(self, prev_data, orient):
        """determine_hand_location (ENV:1286-1307) from the previous episode's last forward pass (None = fresh env)."""
        out = np.zeros(3)
        lib().kgo_hand_location(None if prev_data is None else C.byref(prev_data), int(orient), out.ctypes.data_as(C.POINTER(d)))
        return out

    def env_step(self, e, action):
        """action: 4 values [wrist, f1, f2, f3] or 6 values [slide_x, slide_y, slide_z, f1, f2, f3]."""
        obs = np.zeros(NOBS); rew = d(); info = np.zeros(3)
        a_, a = _p(action)
        fn = lib().kgo_env_step6 if len(a_) == 6 else lib().kgo_env_step
        done = fn(self.h, C.byref(e), a, obs.ctypes.data_as(C.POINTER(d)), C.byref(rew), info.ctypes.data_as(C.POINTER(d)))
        return obs, rew.value, bool(done), info

    def expert_action(self, e, obs, mode):
        act = np.zeros(4); lift = C.c_int()
        _, o = _p(obs)
        lib().kgo_expert_action(C.byref(e), o, mode, act.ctypes.data_as(C.POINTER(d)), C.byref(lift))
        return act, bool(lift.value)
