"""Host-side mirror of the reference env interface on top of the C ABI.

* VecKinovaGripperEnv   -- N environments on one GPU, torch tensors in / out (the product path).
* KinovaGripper_Env     -- single-env facade with the reference's own signatures
                           (gym_kinova_gripper/envs/kinova_gripper_env.py:49, reset :1310, step :1495) so that
                           main_DDPGfD.py / expert_data.py call sites stay unchanged.
* make(id)              -- stand-in for gym.make('gym_kinova_gripper:kinovagripper-v0')
                           (gym_kinova_gripper/__init__.py:3-7: TimeLimit with max_episode_steps=100; every caller
                           then sets env._max_episode_steps = 30, MAIN:164,384, EXP:729).
Everything numerical happens in the CUDA library; this file is shape/orientation/pose bookkeeping."""
import csv
import os

import numpy as np

from .handle import Handle
from .modelcard import ModelCard, SHAPES, ORIENTATIONS

EXPERT_MODES = {"naive": 0, "position-dependent": 1, "combined": 2}


def latin_square_keys(shape_keys, max_elements):
    """Generate_Latin_Square (ENV:895-964): rows of a cyclic Latin square appended until max_elements.
    The env pops from the END of this list (get_object, ENV:988)."""
    keys = list(shape_keys)
    if len(keys) == 0:
        raise ValueError("No shape keys")
    out = []
    while len(out) < max_elements:
        n = len(keys) - 1
        k = n
        done = False
        for _ in range(0, n + 1):
            temp = k
            while temp <= n:
                out.append(keys[temp]); temp += 1
                if len(out) == max_elements:
                    done = True; break
            if done:
                break
            for j in range(0, k):
                out.append(keys[j])
                if len(out) == max_elements:
                    done = True; break
            if done:
                break
            k -= 1
    return out


class VecKinovaGripperEnv:
    """Batched KinovaGripper_Env: env i is an independent episode; all tensors live on `device`."""

    def __init__(self, n_envs, device=0, frame_skip=15, max_episode_steps=30, lib=None, use_torch=True):
        self.n = int(n_envs)
        self.card = ModelCard()
        self.env_offset = 0                           # global id of env 0 (multi-GPU shards, shard.make_sharded_env)
        host = isinstance(device, str) and device == "cpu"   # tests only: the host-emulation library takes host pointers
        if host and lib is None:
            raise ValueError("device='cpu' needs the test-only emulation library; the package has no CPU path")
        self.h = Handle(self.n, device=0 if host else device, lib=lib)
        self.h.set_frame_skip(frame_skip)
        self.h.set_max_episode_steps(max_episode_steps)
        self._max_episode_steps = max_episode_steps
        self.use_torch = use_torch
        if use_torch:
            import torch
            self.torch = torch
            self.device = torch.device("cpu") if host else torch.device("cuda", device)
            f32 = dict(dtype=torch.float32, device=self.device)
            self.obs_soa = torch.zeros((82, self.n), **f32)
            self.reward = torch.zeros(self.n, **f32)
            self.done = torch.zeros(self.n, dtype=torch.int32, device=self.device)
            self.action_soa = torch.zeros((4, self.n), **f32)
            self.action_out = torch.zeros((4, self.n), **f32)
            self.shape_ids = torch.zeros(self.n, dtype=torch.int32, device=self.device)
            if not host:
                self.h.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self.shape_ids_host = np.zeros(self.n, dtype=np.int32)

    # ---- reset
    def set_shapes(self, shape_keys):
        """env i -> shape_keys[i mod len] (the Latin-square order of SURVEY 8d config 2/3)."""
        ids = np.array([SHAPES.index(shape_keys[(self.env_offset + i) % len(shape_keys)]) for i in range(self.n)], dtype=np.int32)
        self.shape_ids_host = ids
        if self.use_torch:
            self.shape_ids.copy_(self.torch.from_numpy(ids))
        return ids

    def reset(self, shape_keys=("CubeS",), hand_orientation="normal", mode="train", variant="A", seed=0, mask=None):
        """Sampled reset of every env (or those with mask != 0).  hand_orientation: 'normal' | 'random' | 'rotated' | 'top'.
        variant 'A' = code-faithful pose tables, 'B' = SURVEY 8d side-grasp variant.  Returns obs [N, 82]."""
        orient = {"normal": 0, "rotated": 1, "top": 2, "random": 3}[hand_orientation]
        if mask is None:
            self.set_shapes(shape_keys)
        m = 0 if mask is None else mask.data_ptr()
        self.h.reset_sampled(m, self.shape_ids.data_ptr(), orient, 1 if variant == "B" else 0, 1 if mode == "test" else 0,
                             seed, self.obs_soa.data_ptr())
        return self.obs_soa.t()

    def reset_explicit(self, shape_ids, hand_euler, obj_pos, mask=None):
        """hand_euler, obj_pos: [N, 3] tensors."""
        t = self.torch
        self.shape_ids.copy_(t.as_tensor(shape_ids, dtype=t.int32, device=self.device))
        e = t.as_tensor(hand_euler, dtype=t.float32, device=self.device).t().contiguous()
        p = t.as_tensor(obj_pos, dtype=t.float32, device=self.device).t().contiguous()
        self.h.reset(0 if mask is None else mask.data_ptr(), self.shape_ids.data_ptr(), e.data_ptr(), p.data_ptr(), self.obs_soa.data_ptr())
        return self.obs_soa.t()

    def enable_global_obs(self):
        """Also produce _get_obs(state_rep="global") (74 values, world frame) at every reset / step; returns the [N, 74]
        view that those calls keep up to date."""
        t = self.torch
        self.obs_global_soa = t.zeros((74, self.n), dtype=t.float32, device=self.device)
        self.h.set_obs_global(self.obs_global_soa.data_ptr())
        return self.obs_global_soa.t()

    # ---- step
    def step(self, actions=None, expert_mode="naive"):
        """actions: [N, 4] tensor ([wrist, f1, f2, f3]) or None for the on-device expert.
        Returns (obs [N,82], reward [N], done [N] bool, info)."""
        a_ptr = 0
        if actions is not None:
            self.action_soa.copy_(actions.t())
            a_ptr = self.action_soa.data_ptr()
        self.h.step(a_ptr, EXPERT_MODES[expert_mode] if isinstance(expert_mode, str) else int(expert_mode),
                    self.obs_soa.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(), self.action_out.data_ptr())
        info = {"lift_reward": self.reward, "finger_reward": 0.0, "grasp_reward": 0.0, "action": self.action_out.t()}
        return self.obs_soa.t(), self.reward, self.done.bool(), info

    def state(self):
        t = self.torch
        qpos = t.zeros((16, self.n), dtype=t.float32, device=self.device)
        qvel = t.zeros((15, self.n), dtype=t.float32, device=self.device)
        self.h.get_state(qpos.data_ptr(), qvel.data_ptr())
        return qpos.t(), qvel.t()


class _Box:
    def __init__(self, low, high):
        self.low, self.high = np.asarray(low, dtype=np.float32), np.asarray(high, dtype=np.float32)
        self.shape = self.low.shape


class KinovaGripper_Env:
    """Single-env facade with the reference's signatures.  Shape / orientation / pose selection follows
    ENV:1149-1283 (Latin-square shape list popped from the end, np.random orientation thresholds, uniformly drawn
    pose row with the 5-character euler truncation); the objects.csv round trip is kept in memory unless
    `persist_objects_csv=True`."""
    metadata = {"render.modes": ["human"]}

    def __init__(self, arm_or_end_effector="hand", frame_skip=15, device=0, lib=None, persist_objects_csv=False):
        if arm_or_end_effector != "hand":
            print("CHOOSE EITHER HAND OR ARM")      # the arm model is out of scope
            raise ValueError
        self.card = ModelCard()
        self.h = Handle(1, device=device, lib=lib)
        self.h.set_frame_skip(frame_skip)
        self.frame_skip = frame_skip
        self._max_episode_steps = 100
        self.h.set_max_episode_steps(10 ** 9)       # TimeLimit is applied by this wrapper (gym semantics)
        self._elapsed_steps = 0
        self.pid = False
        self.with_grasp_reward = False
        self.Grasp_Reward = False
        self.orientation = "normal"
        self.random_shape = "CubeS"
        self.orientation_idx = 0
        self.coords_filename = None
        self.obj_coord_region = None
        self.obj_coords = [0, 0, 0]
        self.objects = {}
        self.obj_keys = []
        self.all_objects = {k: f"/kinova_description/j2s7s300_end_effector_v1_{k}.xml" for k in SHAPES}
        self.action_space = _Box([-0.8] * 4, [0.8] * 4)
        self.Tfw = np.zeros([4, 4])
        self.wrist_pose = np.zeros(3)
        self.persist = persist_objects_csv
        self.env = self                               # gym's TimeLimit exposes the raw env as .env
        self.np_random = np.random

    # ---- small accessors used by MAIN / EXP
    def seed(self, seed=None):
        np.random.seed(seed)
        return [seed]
    def get_obj_coords(self): return self.obj_coords
    def get_orientation(self): return self.orientation
    def get_orientation_idx(self): return self.orientation_idx
    def get_coords_filename(self): return self.coords_filename
    def get_random_shape(self): return self.random_shape
    def set_with_grasp_reward(self, v): self.with_grasp_reward = v
    def get_with_grasp_reward(self): return self.with_grasp_reward
    def get_all_objects(self): return self.all_objects
    def get_obj_keys(self): return self.obj_keys

    def check_obj_file_empty(self, filename):
        if self.persist:
            if not os.path.exists(filename):
                return False
            with open(filename, "r") as f:
                return not f.read(1)
        return len(self.obj_keys) == 0

    def experiment(self, shape_keys):
        for key in shape_keys:
            if key not in self.all_objects:
                print("Invlaid shape key requested")
                raise ValueError
            self.objects[key] = self.all_objects[key]
        if len(shape_keys) == 0:
            print("No shape keys")
            raise ValueError
        return self.objects

    def Generate_Latin_Square(self, max_elements, filename, shape_keys, test=False):
        self.objects = self.experiment(shape_keys)
        self.obj_keys = self.obj_keys + latin_square_keys(list(self.objects.keys()), max_elements)
        if self.persist:
            with open(filename, "w", newline="") as f:
                w = csv.writer(f)
                for key in self.obj_keys:
                    w.writerow(key)      # one character per column, as the reference writes it (ENV:961-964)

    def select_orienation(self, random_shape, hand_orientation):
        orientation_type = 0.330
        if hand_orientation == "random":
            orientation_type = np.random.rand()
        if orientation_type < 0.333:
            return "normal"
        if orientation_type > 0.667:
            return "top"
        return "rotated"

    # ---- reset / step with the reference's signatures
    def reset(self, shape_keys, hand_orientation, with_grasp=False, env_name="env", mode="train", start_pos=None,
              obj_params=None, qpos=None, obj_coord_region=None, with_noise=True):
        self.set_with_grasp_reward(with_grasp)
        self.obj_coord_region = obj_coord_region
        if len(self.objects) == 0:
            self.objects = self.experiment(shape_keys)
        if len(self.obj_keys) == 0:
            self.Generate_Latin_Square(20000 if env_name == "env" else 200, "objects.csv" if env_name == "env" else "eval_objects.csv", shape_keys)
        if obj_params is None:
            random_shape = self.obj_keys.pop()
        else:
            random_shape = obj_params[0] + obj_params[1]
        self.random_shape = random_shape
        self.orientation = self.select_orienation(random_shape, hand_orientation)
        shape_id = SHAPES.index(random_shape)
        tab = self.card.pose("test" if mode == "test" else "train")[shape_id, ORIENTATIONS.index(self.orientation)]
        data = tab[1:]                               # the reference drops file line 0 (ENV:1012)
        if start_pos is None:
            if obj_coord_region is not None and obj_coord_region != "origin":
                rng = {"left": [-.09, -.03], "center": [-.03, .03], "target": [-.01, .01], "right": [.03, .09]}[obj_coord_region]
                cand = np.nonzero((data[:, 0] >= rng[0]) & (data[:, 0] <= rng[1]))[0]
                idx = int(cand[np.random.randint(0, len(cand))])
            else:
                idx = int(np.random.randint(0, len(data)))
            row = data[idx]
            obj = [float(row[0]), float(row[1]), float(row[2])]
            euler = [float(row[6]), float(row[7]), float(row[8])] if with_noise else [0.0, 0.0, 0.0]
            if obj_coord_region == "origin":
                obj, euler, idx = [0.0, 0.0, float(data[0][2])], [0.0, 0.0, 0.0], None
            self.orientation_idx = idx
            self.coords_filename = ("gym_kinova_gripper/envs/kinova_description/obj_hand_coords/" + ("with_noise/" if with_noise else "no_noise/")
                                    + ("test" if mode == "test" else "train") + "_coords/" + self.orientation + "/" + random_shape + ".txt")
        else:
            obj = [float(start_pos[0]), float(start_pos[1]), float(start_pos[2]) if len(start_pos) > 2 else float(self.card.get("mesh_geom").reshape(17, 15)[3 + shape_id, 10])]
            euler = [-1.57, 0.0, -1.57] if self.orientation == "normal" else ([0.0, 0.0, 0.0] if self.orientation == "top" else [-1.2, 0.0, 0.0])
        obs = self.h.reset_host(np.array([shape_id], dtype=np.int32), np.array(euler, dtype=np.float32).reshape(3, 1),
                                np.array(obj, dtype=np.float32).reshape(3, 1))
        self.obj_coords = obj
        self._elapsed_steps = 0
        self.Grasp_Reward = False
        self._last_obs = obs[:, 0]
        self._update_frames(euler)
        states = [float(v) for v in obs[:, 0]]
        if self.pid:
            states = states + [states[81]]
        return states

    def _update_frames(self, euler):
        # Tfw / wrist_pose as _get_trans_mat_wrist_pose leaves them (ENV:274-288); host copy for callers that read env.Tfw
        from .hostmath import hand_frame_from_euler
        self.Tfw, self.wrist_pose = hand_frame_from_euler(self.card, euler, self.h.get_state_host()[0][:3, 0])

    def step(self, action, graspnetwork=False):
        action = np.asarray(action, dtype=np.float32).reshape(-1)
        if len(action) == 6:
            action = action[2:6]                     # [0, 0, wrist, f1, f2, f3] (ENV:1502-1503)
        obs, rew, done = self.h.step_host(action.reshape(4, 1), 0)
        self._elapsed_steps += 1
        lift = float(rew[0])
        d = bool(lift > 0)
        if self._elapsed_steps >= self._max_episode_steps:
            d = True
        self._last_obs = obs[:, 0]
        states = [float(v) for v in obs[:, 0]]
        if self.pid:
            states = states + [states[81]]
        info = {"finger_reward": 0.0, "grasp_reward": 0.0, "lift_reward": lift}
        return states, lift, d, info

    def _get_obs(self, state_rep="local"):
        """ENV:438-534: the observation of the current state, 82 values in the wrist frame ("local") or 74 in the world
        frame ("global"; enabled on first use, valid from the next reset / step on)."""
        if state_rep == "local":
            return [float(v) for v in self._last_obs]
        if state_rep == "global":
            if not getattr(self, "_global_on", False):
                self.h.enable_obs_global()
                self._global_on = True
                raise RuntimeError("global-frame observations are produced from the next reset()/step() on")
            return [float(v) for v in self.h.obs_global_host()[:, 0]]
        raise ValueError(f"unknown state_rep {state_rep!r}")

    def _get_dot_product(self, obj_state=None):
        return float(self._last_obs[81])

    def close(self):
        self.h.close()


def make(env_id="gym_kinova_gripper:kinovagripper-v0", **kw):
    if not env_id.endswith("kinovagripper-v0"):
        raise ValueError(f"unknown env id {env_id}")
    return KinovaGripper_Env(arm_or_end_effector="hand", frame_skip=15, **kw)
