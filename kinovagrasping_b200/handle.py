"""Thin object wrapper over the C ABI (include/kg.h): error handling + numpy host-buffer helpers.
Device-pointer entry points take raw integer addresses (e.g. torch.Tensor.data_ptr())."""
import ctypes as C

import numpy as np

from . import _lib
from .modelcard import DEFAULT_PATH


class KgError(RuntimeError):
    pass


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Handle:
    """One batch of environments on one device (kg_create ... kg_destroy)."""

    def __init__(self, n_envs, device=0, blob_path=DEFAULT_PATH, lib=None):
        self.lib = lib if lib is not None else _lib.load()
        with open(blob_path, "rb") as f:
            blob = f.read()
        self.n = int(n_envs)
        self.h = C.c_void_p()
        rc = self.lib.kg_create(self.n, int(device), blob, len(blob), C.byref(self.h))
        if rc != 0:
            msg = self.lib.kg_last_error(self.h).decode() if self.h else "kg_create failed"
            raise KgError(f"kg_create: {msg} (rc={rc})")

    def _chk(self, rc, what):
        if rc != 0:
            raise KgError(f"{what}: {self.lib.kg_last_error(self.h).decode()} (rc={rc})")

    def close(self):
        if self.h:
            self.lib.kg_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- device-pointer API (addresses as ints; 0/None = NULL)
    def set_stream(self, stream_ptr):
        self._chk(self.lib.kg_set_stream(self.h, C.c_void_p(stream_ptr or 0)), "kg_set_stream")

    def sync(self):
        self._chk(self.lib.kg_sync(self.h), "kg_sync")

    def reset(self, mask, shape, euler, objpos, obs):
        self._chk(self.lib.kg_reset(self.h, mask or None, shape, euler, objpos, obs or None), "kg_reset")

    def reset_pose(self, mask, shape, euler, objpos, handpos, obs):
        self._chk(self.lib.kg_reset_pose(self.h, mask or None, shape, euler, objpos, handpos or None, obs or None), "kg_reset_pose")

    def set_hand_location(self, mode):
        """0 = slides start at 0 (a new env per episode, MAIN:381); 1 = determine_hand_location with the previous episode's Tfw (EXP:696)."""
        self._chk(self.lib.kg_set_hand_location(self.h, int(mode)), "kg_set_hand_location")

    def set_shape_cycle(self, shape_ids):
        """Latin-square shape schedule for sampled resets with no shape array: episode e of env i -> shape_ids[(i + e) % len]."""
        a = np.ascontiguousarray(shape_ids, dtype=np.int32).reshape(-1)
        self._chk(self.lib.kg_set_shape_cycle(self.h, _np_ptr(a) if a.size else None, int(a.size)), "kg_set_shape_cycle")

    def load_grasp_forest(self, forest_path=None):
        """Upload the grasp classifier's trees (assets/gc_forest.bin, tools/extract_forest.py; ENV:248) once."""
        if not getattr(self, "_forest_loaded", False):
            import os
            path = forest_path or os.path.join(os.path.dirname(DEFAULT_PATH), "gc_forest.bin")
            with open(path, "rb") as f:
                blob = f.read()
            self._chk(self.lib.kg_set_grasp_forest(self.h, blob, len(blob)), "kg_set_grasp_forest")
            self._forest_loaded = True

    def set_grasp_reward(self, on, forest_path=None):
        """set_with_grasp_reward (ENV:791): the grasp classifier's +5.0 (ENV:643-659) is added by every step."""
        if on:
            self.load_grasp_forest(forest_path)
        self._chk(self.lib.kg_set_grasp_reward(self.h, 1 if on else 0), "kg_set_grasp_reward")

    def grasp_predict_host(self, x):
        """Grasp_net.predict(x) (ENV:651) by the step's own forest kernels; x: [m, n_features], m <= n_envs."""
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float32).T)
        out = np.zeros(x.shape[1], dtype=np.float32)
        self._chk(self.lib.kg_grasp_predict_host(self.h, _np_ptr(x), int(x.shape[1]), _np_ptr(out)), "kg_grasp_predict_host")
        return out

    def get_grasp_reward_host(self):
        """(info["grasp_reward"] of the last step float[n], Grasp_Reward flags bool[n])"""
        g = np.zeros(self.n, dtype=np.float32)
        f = np.zeros(self.n, dtype=np.int32)
        self._chk(self.lib.kg_get_grasp_reward_host(self.h, _np_ptr(g), _np_ptr(f)), "kg_get_grasp_reward_host")
        return g, f != 0

    def orientation_host(self):
        """Hand orientation each env's last sampled reset drew: 0 normal, 1 rotated, 2 top (-1: explicit-pose reset)."""
        o = np.zeros(self.n, dtype=np.int32)
        self._chk(self.lib.kg_get_orientation_host(self.h, _np_ptr(o)), "kg_get_orientation_host")
        return o

    def set_action_dim(self, dim):
        """4 = [wrist, f1, f2, f3]; 6 = [slide_x, slide_y, slide_z, f1, f2, f3] (ENV:1502-1503)."""
        self._chk(self.lib.kg_set_action_dim(self.h, int(dim)), "kg_set_action_dim")
        self.act_dim = int(dim)

    def reset_sampled(self, mask, shape, orient, variant, test_table, seed, obs):
        self._chk(self.lib.kg_reset_sampled(self.h, mask or None, shape or None, int(orient), int(variant), int(test_table),
                                            C.c_uint64(int(seed) & (2 ** 64 - 1)), obs or None), "kg_reset_sampled")

    def step(self, action, expert_mode, obs, reward, done, action_out=None):
        self._chk(self.lib.kg_step(self.h, action or None, int(expert_mode), obs or None, reward or None, done or None,
                                   action_out or None), "kg_step")

    def expert_action(self, expert_mode, action):
        """controller.get_action(obs) (EXP:610) alone, into the device buffer `action` ([4][n] / [6][n]); pass it to step()."""
        self._chk(self.lib.kg_expert_action(self.h, int(expert_mode), action), "kg_expert_action")

    def expert_action_host(self, expert_mode=0, out=None):
        a = np.zeros((getattr(self, "act_dim", 4), self.n), dtype=np.float32) if out is None else out
        self._chk(self.lib.kg_expert_action_host(self.h, int(expert_mode), _np_ptr(a)), "kg_expert_action_host")
        return a

    def get_state(self, qpos, qvel):
        self._chk(self.lib.kg_get_state(self.h, qpos or None, qvel or None), "kg_get_state")

    def set_state(self, qpos, qvel):
        self._chk(self.lib.kg_set_state(self.h, qpos or None, qvel or None), "kg_set_state")

    def set_max_episode_steps(self, n):
        self._chk(self.lib.kg_set_max_episode_steps(self.h, int(n)), "kg_set_max_episode_steps")

    def set_frame_skip(self, n):
        self._chk(self.lib.kg_set_frame_skip(self.h, int(n)), "kg_set_frame_skip")

    def set_env_offset(self, off):
        self._chk(self.lib.kg_set_env_offset(self.h, int(off)), "kg_set_env_offset")

    def set_schedule(self, mode):
        self._chk(self.lib.kg_set_schedule(self.h, int(mode)), "kg_set_schedule")

    def set_mode(self, mode):
        """0 automatic, 1 always split (light workspaces + heavy lane), 2 always full workspaces."""
        self._chk(self.lib.kg_set_mode(self.h, int(mode)), "kg_set_mode")

    def set_obs_global(self, ptr):
        """float[74][n] buffer (device pointer; 0 = off) that every step / reset also fills with the global-frame obs."""
        self._chk(self.lib.kg_set_obs_global(self.h, C.c_void_p(ptr or 0)), "kg_set_obs_global")

    def enable_obs_global(self):
        """Route the global-frame obs of every step / reset into the handle's own buffer (read it with obs_global_host)."""
        buf = C.c_void_p()
        self._chk(self.lib.kg_obs_global_buffer(self.h, C.byref(buf)), "kg_obs_global_buffer")
        self.set_obs_global(buf.value)
        return buf.value

    def obs_global_host(self):
        out = np.zeros((74, self.n), dtype=np.float32)
        self._chk(self.lib.kg_get_obs_global_host(self.h, _np_ptr(out)), "kg_get_obs_global_host")
        return out

    def set_pipeline(self, groups):
        """Number of env groups (1..4) whose kernel chains kg_step overlaps on separate internal streams."""
        self._chk(self.lib.kg_set_pipeline(self.h, int(groups)), "kg_set_pipeline")

    def measure_fp32_peak(self):
        """TFLOP/s of an FP32 FMA microbenchmark on this device (the measured FP32 roofline denominator)."""
        v = C.c_float()
        self._chk(self.lib.kg_measure_fp32_peak(self.h, C.byref(v)), "kg_measure_fp32_peak")
        return float(v.value)

    def set_graph(self, on):
        """1 (default) = kg_step replays a captured CUDA graph of its kernel chain; 0 = individual launches."""
        self._chk(self.lib.kg_set_graph(self.h, int(on)), "kg_set_graph")

    def set_dependent_launch(self, on):
        """1 = programmatic dependent launch between the kernels of a group's chain (launch latency hidden); results unchanged."""
        self._chk(self.lib.kg_set_dependent_launch(self.h, int(on)), "kg_set_dependent_launch")

    def set_timing(self, on):
        self._chk(self.lib.kg_set_timing(self.h, int(on)), "kg_set_timing")

    def timing_breakdown(self):
        """ms of the last step by launch kind (timing on): schedule, begin, narrowphase, substep, last, rays."""
        b = (C.c_float * 8)()
        self._chk(self.lib.kg_timing_breakdown(self.h, b), "kg_timing_breakdown")
        return {"schedule": b[1], "begin": b[2], "narrowphase": b[3], "substep": b[4], "last": b[5], "rays": b[6]}

    def last_action_host(self):
        a = np.zeros((4, self.n), dtype=np.float32)
        self._chk(self.lib.kg_get_last_action_host(self.h, _np_ptr(a)), "kg_get_last_action_host")
        return a

    def last_step_ms(self):
        ms = C.c_float()
        self._chk(self.lib.kg_last_step_ms(self.h, C.byref(ms)), "kg_last_step_ms")
        return ms.value

    def stats(self):
        out = (C.c_longlong * 8)()
        self._chk(self.lib.kg_get_stats(self.h, out), "kg_get_stats")
        return {"mpr_pairs": out[0], "newton_iters": out[1], "contact_overflow": out[2], "substeps": out[3], "contacts": out[4]}

    # ---- host-buffer API (numpy)
    def reset_host(self, shape, euler, objpos, mask=None, handpos=None):
        n = self.n
        shape = np.ascontiguousarray(shape, dtype=np.int32).reshape(n)
        euler = np.ascontiguousarray(euler, dtype=np.float32).reshape(3, n)
        objpos = np.ascontiguousarray(objpos, dtype=np.float32).reshape(3, n)
        obs = np.zeros((82, n), dtype=np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.int32).reshape(n)
        hp = None if handpos is None else np.ascontiguousarray(handpos, dtype=np.float32).reshape(3, n)
        self._chk(self.lib.kg_reset_pose_host(self.h, None if m is None else _np_ptr(m), _np_ptr(shape), _np_ptr(euler), _np_ptr(objpos),
                                              None if hp is None else _np_ptr(hp), _np_ptr(obs)), "kg_reset_pose_host")
        return obs

    def step_host(self, action=None, expert_mode=0, obs_out=None, reward_out=None, done_out=None):
        n = self.n
        a = None if action is None else np.ascontiguousarray(action, dtype=np.float32).reshape(getattr(self, "act_dim", 4), n)
        obs = np.zeros((82, n), dtype=np.float32) if obs_out is None else obs_out
        rew = np.zeros(n, dtype=np.float32) if reward_out is None else reward_out
        done = np.zeros(n, dtype=np.int32) if done_out is None else done_out
        self._chk(self.lib.kg_step_host(self.h, None if a is None else _np_ptr(a), int(expert_mode), _np_ptr(obs), _np_ptr(rew),
                                        _np_ptr(done)), "kg_step_host")
        return obs, rew, done

    def get_state_host(self):
        qpos = np.zeros((16, self.n), dtype=np.float32)
        qvel = np.zeros((15, self.n), dtype=np.float32)
        self._chk(self.lib.kg_get_state_host(self.h, _np_ptr(qpos), _np_ptr(qvel)), "kg_get_state_host")
        return qpos, qvel

    def set_state_host(self, qpos=None, qvel=None):
        qp = None if qpos is None else np.ascontiguousarray(qpos, dtype=np.float32).reshape(16, self.n)
        qv = None if qvel is None else np.ascontiguousarray(qvel, dtype=np.float32).reshape(15, self.n)
        self._chk(self.lib.kg_set_state_host(self.h, None if qp is None else _np_ptr(qp), None if qv is None else _np_ptr(qv)),
                  "kg_set_state_host")

    def forward_debug(self, env, ctrl):
        c = np.ascontiguousarray(ctrl, dtype=np.float32).reshape(9)
        out = np.zeros(_lib.DEBUG_FLOATS, dtype=np.float32)
        self._chk(self.lib.kg_forward_debug(self.h, int(env), _np_ptr(c), _np_ptr(out)), "kg_forward_debug")
        ncon = int(out[300])
        con = out[301:301 + 8 * 23].reshape(23, 8)
        return {"M": out[:225].reshape(15, 15), "bias": out[225:240], "fsmooth": out[240:255], "qacc_smooth": out[255:270],
                "qacc": out[270:285], "ftotal": out[285:300], "ncon": ncon, "con": con[:min(ncon, 23)],
                "geom_xpos": out[485:512].reshape(9, 3), "geom_xmat": out[512:593].reshape(9, 9),
                "ngroup": int(out[593]), "groups": out[594:850].reshape(32, 8), "Jbase": out[850:2020].reshape(78, 15)}

    @staticmethod
    def explicit_rows(dbg):
        """The scalar constraint rows of a forward_debug dump, as (kind, J[15], aref, D) in the kernel's group order:
        kind 1 tendon equality, 2 joint limit, 3 pyramidal contact (four rows per contact: n +- mu t1, n +- mu t2)."""
        rows = []
        for g in range(min(dbg["ngroup"], 32)):
            tp, D, mu = int(dbg["groups"][g, 0]), float(dbg["groups"][g, 1]), float(dbg["groups"][g, 2])
            b = int(dbg["groups"][g, 7])
            if tp in (1, 2):
                rows.append((tp, dbg["Jbase"][b].astype(np.float64), float(dbg["groups"][g, 3]), D))
            elif tp == 3:
                n, t1, t2 = (dbg["Jbase"][b + k].astype(np.float64) for k in range(3))
                for k, J in enumerate((n + mu * t1, n - mu * t1, n + mu * t2, n - mu * t2)):
                    rows.append((3, J, float(dbg["groups"][g, 3 + k]), D))
        return rows
