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

    def set_shape_cycle(self, shape_keys):
        """Latin-square schedule on the device (ENV:895-964's property over (env, episode)): from now on a sampled reset that
        is given no shapes (`reset(mask=...)`, the auto-reset of finished envs) moves env i on to the next key of its row,
        shape_keys[(i + episode) % len].  An empty list switches it off (envs keep their shape)."""
        self.h.set_shape_cycle([SHAPES.index(k) for k in shape_keys])
        self._cycle = len(shape_keys) > 0

    def reset(self, shape_keys=("CubeS",), hand_orientation="normal", mode="train", variant="A", seed=0, mask=None):
        """Sampled reset of every env (or those with mask != 0).  hand_orientation: 'normal' | 'random' | 'rotated' | 'top'.
        variant 'A' = code-faithful pose tables, 'B' = SURVEY 8d side-grasp variant.  Returns obs [N, 82]."""
        orient = {"normal": 0, "rotated": 1, "top": 2, "random": 3}[hand_orientation]
        if mask is None:
            self.set_shapes(shape_keys)
        m = 0 if mask is None else mask.data_ptr()
        shapes = 0 if (mask is not None and getattr(self, "_cycle", False)) else self.shape_ids.data_ptr()
        self.h.reset_sampled(m, shapes, orient, 1 if variant == "B" else 0, 1 if mode == "test" else 0,
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
        if getattr(self, "with_grasp_reward", False):      # reward = lift (0 | 50) + grasp (0 | 5), both exact in fp32 (ENV:682)
            lift = (self.reward >= 50.0).to(self.reward.dtype) * 50.0
            info = {"lift_reward": lift, "finger_reward": 0.0, "grasp_reward": self.reward - lift, "action": self.action_out.t()}
        else:
            info = {"lift_reward": self.reward, "finger_reward": 0.0, "grasp_reward": 0.0, "action": self.action_out.t()}
        return self.obs_soa.t(), self.reward, self.done.bool(), info

    def set_with_grasp_reward(self, with_grasp):
        """ENV:791 for the whole batch: the grasp classifier's +5.0 (ENV:643-659, on-device decision forest) is added to the
        reward of the first step of each episode whose obs[0:75] it labels a grasp."""
        self.h.set_grasp_reward(bool(with_grasp))
        self.with_grasp_reward = bool(with_grasp)

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


# The reference keeps two pieces of state in FILES that outlive an env object (MAIN:381 builds a new env per episode and
# relies on them): the remaining Latin-square shape list in ./objects.csv / ./eval_objects.csv (ENV:961-999) and the hand
# euler inside each shape's XML (ENV:851-877).  Unless persist_objects_csv=True, the lists live here, keyed by filename.
_MEM_FILES = {}
# euler of body j2s7s300_link_7 as the shipped XMLs hold it (XML:59 of each file); write_xml overwrites it
_XML_EULER = {k: [-1.57, 0.0, -1.57] for k in SHAPES}
_XML_EULER["CubeS"] = [-0.06, -0.16, -0.13]
_XML_EULER.update({"LemonS": [-0.09, -0.16, -0.15], "LemonM": [-1.22, -0.04, 0.029], "LemonB": [-0.10, -0.17, -0.02]})
_XML_FILE = {"LemonS": "slemon", "LemonM": "mlemon", "LemonB": "blemon"}      # ENV:198-201
for _k in ("CubeM", "CylinderM", "Cube45M", "Cone2M", "Vase1M", "Vase2M"):      # the committed XMLs of the medium sizes (Cone1M has the side pose)
    _XML_EULER[_k] = [0.0, 0.0, 0.0]


def truncate_euler(rot):
    """write_xml (ENV:870-874): every component goes into the XML as the first 5 characters of str(value)."""
    out = []
    for v in rot:
        t = str(np.float64(v))[0:5]
        try:
            out.append(float(t))
        except ValueError:          # e.g. '-8.7e' -- the XML would not load in the reference either
            raise ValueError(f"hand euler {v!r} does not survive the reference's 5-character truncation")
    return out


class KinovaGripper_Env:
    """Single-env facade with the reference's signatures; also stands in for gym's TimeLimit wrapper (`.env` is self).
    Shape / orientation / pose selection restates ENV:884-1054 and ENV:1149-1410 call for call (same np.random draws in
    the same order, the same list/"file" bookkeeping and the same quirks -- each noted where it happens), so
    main_DDPGfD.py / expert_data.py run unchanged on it (tests/test_callsite.py)."""
    metadata = {"render.modes": ["human"]}

    def __init__(self, arm_or_end_effector="hand", frame_skip=15, device=0, lib=None, persist_objects_csv=False, handle=None):
        if arm_or_end_effector != "hand":
            print("CHOOSE EITHER HAND OR ARM")      # the arm model is out of scope
            raise ValueError
        self.card = ModelCard()
        self.h = handle if handle is not None else Handle(1, device=device, lib=lib)
        self.h.set_frame_skip(frame_skip)
        self.h.set_action_dim(6)                    # step() pads 4-d actions itself, as ENV:1502-1503 does
        self.frame_skip = frame_skip
        self._max_episode_steps = 100               # gym registration (gym_kinova_gripper/__init__.py:3-7)
        self.h.set_max_episode_steps(10 ** 9)       # TimeLimit is applied by this wrapper (gym semantics)
        self._elapsed_steps = 0
        self.pid = False
        self.state_rep = "local"
        self._with_grasp_reward = False
        self.Grasp_Reward = False
        self.obj_size = "m"                         # ENV:62; never changed on the Latin-square path
        self.filename = "/kinova_description/j2s7s300_end_effector_v1_mbox.xml"
        self.orientation = "normal"
        self.random_shape = "CubeS"
        self.orientation_idx = 0
        self.coords_filename = None
        self.obj_coord_region = None
        self.obj_coords = [0, 0, 0]
        self.objects = {}
        self.obj_keys = []
        self.all_objects = {k: f"/kinova_description/j2s7s300_end_effector_v1_{_XML_FILE.get(k, k)}.xml" for k in SHAPES}
        self.action_space = _Box([-0.8] * 4, [0.8] * 4)
        self.Tfw = np.zeros([4, 4])                 # ENV:110: a new env has no hand frame yet
        self.wrist_pose = np.zeros(3)
        self.persist = persist_objects_csv
        self.env = self                               # gym's TimeLimit exposes the raw env as .env
        self.np_random = np.random
        self._shape_id = 0
        self._euler = list(_XML_EULER["CubeS"])
        self._last_obs = np.zeros(82, dtype=np.float32)

    # ---- small accessors used by MAIN / EXP (ENV:748-805)
    def seed(self, seed=None):
        np.random.seed(seed)
        return [seed]
    def set_obj_coords(self, x, y, z): self.obj_coords[0] = x; self.obj_coords[1] = y; self.obj_coords[2] = z
    def get_obj_coords(self): return self.obj_coords
    def set_random_shape(self, shape): self.random_shape = shape
    def get_random_shape(self): return self.random_shape
    def set_orientation_idx(self, idx): self.orientation_idx = idx
    def get_orientation_idx(self): return self.orientation_idx
    def set_obj_coord_region(self, region): self.obj_coord_region = region
    def get_obj_coord_region(self): return self.obj_coord_region
    def get_orientation(self): return self.orientation
    def set_orientation(self, orientation): self.orientation = orientation
    def get_with_grasp_reward(self): return self.with_grasp_reward
    def set_with_grasp_reward(self, with_grasp):
        """ENV:791.  ENV:643-659 adds +5 once per episode when the sklearn forest of KD/gc_model.pkl labels local_obs[0:75]
        >= 0.3; here the forest is a device table walked by two small kernels behind the step (csrc/kg_forest.cuh)."""
        self.with_grasp_reward = with_grasp
    # MAIN:283,302 assign env.with_grasp_reward directly, so the attribute itself switches the kernels
    @property
    def with_grasp_reward(self): return self._with_grasp_reward
    @with_grasp_reward.setter
    def with_grasp_reward(self, with_grasp):
        on = with_grasp is True          # `if with_grasp_reward is True` (ENV:643): only the bool True enables it
        if on != (self._with_grasp_reward is True):
            self.h.set_grasp_reward(on)
        self._with_grasp_reward = with_grasp
    def get_coords_filename(self): return self.coords_filename
    def set_coords_filename(self, coords_filename): self.coords_filename = coords_filename
    def get_all_objects(self): return self.all_objects
    def get_obj_keys(self): return self.obj_keys

    # ---- the shape-list "files" (ENV:884-893, 961-999)
    def _file_exists(self, filename):
        return os.path.exists(filename) if self.persist else filename in _MEM_FILES

    def _file_write(self, filename, keys):
        if self.persist:
            with open(filename, "w", newline="") as f:
                w = csv.writer(f)
                for key in keys:
                    w.writerow(key)      # one character per column, as the reference writes it (ENV:961-964)
        else:
            _MEM_FILES[filename] = list(keys)

    def _file_read(self, filename):
        if self.persist:
            with open(filename, newline="") as f:
                return ["".join(row) for row in csv.reader(f)]
        return list(_MEM_FILES[filename])

    def check_obj_file_empty(self, filename):
        """ENV:884-893: a MISSING file counts as not empty."""
        if not self._file_exists(filename):
            return False
        if self.persist:
            with open(filename, "r") as f:
                return not f.read(1)
        return len(_MEM_FILES[filename]) == 0

    def experiment(self, shape_keys):
        """ENV:807-818: keys ACCUMULATE in self.objects over calls; a mismatch with the request is an error."""
        for key in shape_keys:
            self.objects[key] = self.all_objects[key]     # KeyError for a shape outside the set, as in the reference
        if len(shape_keys) == 0:
            print("No shape keys")
            raise ValueError
        elif len(shape_keys) != len(self.objects):
            print("Invlaid shape key requested")
            raise ValueError
        return self.objects

    def Generate_Latin_Square(self, max_elements, filename, shape_keys, test=False):
        """ENV:895-964: APPENDS max_elements keys to self.obj_keys, then writes the whole list to the file."""
        self.objects = self.experiment(shape_keys)
        self.obj_keys = self.obj_keys + latin_square_keys(list(self.objects.keys()), max_elements)
        self._file_write(filename, self.obj_keys)

    def objects_file_to_list(self, filename, num_objects, shape_keys):
        """ENV:966-981.  Quirk kept: when the file has to be generated the keys end up in self.obj_keys TWICE (once
        appended by Generate_Latin_Square, once more by reading the file back)."""
        if self._file_exists(filename):
            if self.check_obj_file_empty(filename):
                print("Object file is empty!")
                self.Generate_Latin_Square(num_objects, filename, shape_keys)
        else:
            self.Generate_Latin_Square(num_objects, filename, shape_keys)
        for row in self._file_read(filename):
            self.obj_keys.append(row)

    def get_object(self, filename):
        """ENV:986-1005: pop from the END, write the remaining list back."""
        random_shape = self.obj_keys.pop()
        self._file_write(filename, self.obj_keys)
        return random_shape, self.objects[random_shape]

    def obj_shape_generator(self, obj_params):
        """ENV:1056-1146 for the single-geom shapes of SURVEY 2 (obj_params = [shape_name, size])."""
        key = obj_params[0] + obj_params[1]
        if key not in self.all_objects:
            raise ValueError(f"object {key!r} is outside the CubeS..Vase2B set")
        self.filename = self.all_objects[key]
        return key

    def select_object(self, env_name, shape_keys, obj_params):
        """ENV:1149-1177."""
        if env_name == "env":
            obj_list_filename, num_objects = "objects.csv", 20000
        else:
            obj_list_filename, num_objects = "eval_objects.csv", 200
        if len(self.objects) == 0:
            self.objects = self.experiment(shape_keys)
        if len(self.obj_keys) == 0:
            self.objects_file_to_list(obj_list_filename, num_objects, shape_keys)
        if obj_params is None:
            random_shape, self.filename = self.get_object(obj_list_filename)
        else:
            random_shape = self.obj_shape_generator(obj_params)
        return random_shape

    def select_orienation(self, random_shape, hand_orientation):
        """ENV:1180-1222 (same np.random calls, including the RBowl / Lemon branches)."""
        orientation_type = 0.330
        if random_shape.find("RBowl") != -1:
            if hand_orientation == "random":
                orientation_type = np.random.uniform(0.333, 1)
        elif random_shape.find("Lemon") != -1:
            if hand_orientation == "random":
                Choice1 = np.random.uniform(0, 0.333)
                Choice2 = np.random.uniform(0.667, 1)
                orientation_type = np.random.choice([Choice1, Choice2])
        elif hand_orientation == "random":
            orientation_type = np.random.rand()
        if orientation_type < 0.333:
            return "normal"
        if orientation_type > 0.667:
            return "top"
        return "rotated"

    def _get_obj_size(self):
        """ENV:706-746 for a single-geom object: the mesh geom's box half sizes with the axis re-ordering heuristic."""
        size = [float(v) for v in self.card.get("mesh_geom").reshape(-1, 15)[3 + self._shape_id, 8:11]]
        if size[2] == 0:
            size[2] = size[1]; size[1] = size[0]
        diffs = [abs(size[0] - size[1]), abs(size[1] - size[2]), abs(size[0] - size[2])]
        if int(np.argmin(diffs)) != 0:
            size[0], size[2] = size[2], size[0]
        return [max(size[0], 0), max(size[1], 0), 0 + size[2]]

    def determine_hand_location(self):
        """ENV:1286-1307.  Evaluated BEFORE the new hand euler is written, so self.Tfw is still the previous episode's
        (all zeros for a freshly constructed env, ENV:110): the quirk SURVEY 8a documents."""
        if self.orientation == "normal":
            return 0, 0, 0, 0, 0, 0
        if self.orientation == "top":
            Z = {"b": 0.15, "m": 0.14, "s": 0.13}[self.obj_size]
            stuff = np.matmul(self.Tfw[0:3, 0:3], [-0.005, -0.155, Z + 0.06])
            return -stuff[0], -stuff[1], stuff[2], 0, 0, 0
        temp = np.matmul(self.Tfw[0:3, 0:3], np.array([0.051, -0.075, 0.06]))
        return -temp[0], -temp[1], temp[2], 0, 0, 0

    def sample_initial_object_hand_pos(self, coords_filename, with_noise=True, orient_idx=None, region=None):
        """ENV:1008-1054 on the model card's copy of the coordinate file (file line 0 is consumed by the delimiter
        sniffing, ENV:1012, so data[i] is file line i + 1).  Quirk kept: with a region, the index is drawn over the
        FILTERED list but then applied to the unfiltered one (ENV:1040-1046)."""
        mode = "test" if "/test_coords/" in coords_filename else "train"
        tab = self.card.pose(mode)[self._shape_id, ORIENTATIONS.index(self.orientation)]
        data = [[float(r[0]), float(r[1]), float(r[2])] + ([float(r[3]), float(r[4]), float(r[5])] if with_noise is True else [0, 0, 0]) for r in tab[1:]]
        if orient_idx is None:
            if region is not None:
                all_regions = {"left": [-.09, -.03], "center": [-.03, .03], "target": [-.01, .01], "right": [.03, .09], "origin": [0, 0]}
                if region == "origin":
                    return 0, 0, data[0][2], 0, 0, 0, None
                sampling_range = all_regions[region]
                region_data = [d for d in data if sampling_range[0] <= d[0] <= sampling_range[1]]
                orient_idx = np.random.randint(0, len(region_data))
            else:
                orient_idx = np.random.randint(0, len(data))
        c = data[orient_idx]
        return c[0], c[1], c[2], c[3], c[4], c[5], orient_idx

    def write_xml(self, new_rotation):
        """ENV:851-881 without the file system: the hand euler the shape's XML now holds (5-character truncation)."""
        self._euler = truncate_euler(new_rotation)
        _XML_EULER[self.random_shape] = list(self._euler)

    def determine_obj_hand_coords(self, random_shape, mode, with_noise=True):
        """ENV:1225-1283."""
        if with_noise is not True:
            # the no_noise directories are capitalised ('Normal/') while the path is built in lower case (ENV:1245): on a
            # case-sensitive file system the reference fails right here
            raise FileNotFoundError("obj_hand_coords/no_noise/<mode>_coords/" + str(self.orientation) + "/" + random_shape + ".txt")
        if mode != "test":
            mode = "train"
        coords_filename = ("gym_kinova_gripper/envs/kinova_description/obj_hand_coords/with_noise/" + str(mode) + "_coords/"
                           + str(self.orientation) + "/" + random_shape + ".txt")
        obj_x, obj_y, obj_z, hand_x, hand_y, hand_z, orient_idx = self.sample_initial_object_hand_pos(
            coords_filename, with_noise=with_noise, orient_idx=None, region=self.obj_coord_region)
        self.write_xml(np.array([hand_x, hand_y, hand_z]))
        return obj_x, obj_y, obj_z, hand_x, hand_y, hand_z, orient_idx, coords_filename

    # ---- reset / step with the reference's signatures
    def reset(self, shape_keys, hand_orientation, with_grasp=False, env_name="env", mode="train", start_pos=None,
              obj_params=None, qpos=None, obj_coord_region=None, with_noise=True):
        """ENV:1310-1410."""
        self.set_with_grasp_reward(with_grasp)
        self.set_obj_coord_region(obj_coord_region)
        random_shape = self.select_object(env_name, shape_keys, obj_params)
        self.set_random_shape(random_shape)
        self._shape_id = SHAPES.index(random_shape)
        self._euler = list(_XML_EULER[random_shape])       # load_model_from_path of that shape's XML (ENV:1002)
        self.set_orientation(self.select_orienation(random_shape, hand_orientation))
        xloc, yloc, zloc, f1prox, f2prox, f3prox = self.determine_hand_location()
        if qpos is not None:
            raise NotImplementedError("reset(qpos=...) (ENV:349-353) is a test hook outside the hot path")
        if start_pos is None:
            obj_x, obj_y, obj_z, hand_x, hand_y, hand_z, orient_idx, coords_filename = self.determine_obj_hand_coords(random_shape, mode, with_noise=with_noise)
            self.set_orientation_idx(orient_idx)
            self.set_coords_filename(coords_filename)
        elif len(start_pos) == 3:
            obj_x, obj_y, obj_z = start_pos[0], start_pos[1], start_pos[2]
        elif len(start_pos) == 2:
            obj_x, obj_y = start_pos[0], start_pos[1]
            obj_z = self._get_obj_size()[-1]
        else:
            xloc, yloc, zloc, f1prox, f2prox, f3prox = start_pos[0], start_pos[1], start_pos[2], start_pos[3], start_pos[4], start_pos[5]
            obj_x, obj_y, obj_z = start_pos[6], start_pos[7], start_pos[8]
        if f1prox != 0 or f2prox != 0 or f3prox != 0:
            raise NotImplementedError("non-zero proximal start angles (9-value start_pos) are outside the hot path")
        self.Grasp_Reward = False
        obs = self.h.reset_host(np.array([self._shape_id], dtype=np.int32), np.array(self._euler, dtype=np.float32).reshape(3, 1),
                                np.array([obj_x, obj_y, obj_z], dtype=np.float32).reshape(3, 1),
                                handpos=np.array([xloc, yloc, zloc], dtype=np.float32).reshape(3, 1))
        self._elapsed_steps = 0
        self._last_obs = obs[:, 0]
        self.set_obj_coords(obj_x, obj_y, obj_z)
        self._update_frames()
        states = [float(v) for v in obs[:, 0]]
        if self.pid:
            states = states + [states[81]]      # ENV:530-531
        return states

    def _update_frames(self):
        # Tfw / wrist_pose as _get_trans_mat_wrist_pose leaves them (ENV:274-288); host copy for callers that read env.Tfw
        from .hostmath import hand_frame_from_euler
        self.Tfw, self.wrist_pose = hand_frame_from_euler(self.card, self._euler, self.h.get_state_host()[0][:3, 0])

    def step(self, action, graspnetwork=False):
        """ENV:1495-1552 under gym's TimeLimit."""
        action = [float(v) for v in np.asarray(action, dtype=np.float64).reshape(-1)]
        if len(action) == 4:
            action = [0, 0, action[0], action[1], action[2], action[3]]      # ENV:1502-1503
        obs, rew, done = self.h.step_host(np.array(action[0:6], dtype=np.float32).reshape(6, 1), 0)
        self._elapsed_steps += 1
        grasp = 0.0
        if self._with_grasp_reward is True:
            g, given = self.h.get_grasp_reward_host()
            grasp, self.Grasp_Reward = float(g[0]), bool(given[0])
        lift = float(rew[0]) - grasp
        d = bool(lift > 0)
        if self._elapsed_steps >= self._max_episode_steps:
            d = True
        self._last_obs = obs[:, 0]
        self._update_frames()
        states = [float(v) for v in obs[:, 0]]
        if self.pid:
            states = states + [states[81]]
        info = {"finger_reward": 0.0, "grasp_reward": grasp, "lift_reward": lift}
        return states, 0.2 * 0.0 + lift + grasp, d, info      # ENV:682

    def _get_obs(self, state_rep=None):
        """ENV:438-534: the observation of the current state, 82 values in the wrist frame ("local") or 74 in the world
        frame ("global"; enabled on first use, valid from the next reset / step on)."""
        if state_rep is None:
            state_rep = self.state_rep
        if state_rep == "local":
            return [float(v) for v in self._last_obs]
        if state_rep == "global":
            if not getattr(self, "_global_on", False):
                self.h.enable_obs_global()
                self._global_on = True
                raise RuntimeError("global-frame observations are produced from the next reset()/step() on")
            return [float(v) for v in self.h.obs_global_host()[:, 0]]
        raise ValueError(f"unknown state_rep {state_rep!r}")

    def _get_obj_pose(self):
        """ENV:564-566: world position of the object geom (entries 21:24 of the global layout)."""
        return np.array(self._get_obs("global")[21:24])

    def _get_dot_product(self, obj_state=None):
        return float(self._last_obs[81])

    def render_img(self, *a, **k):
        return None                                  # rendering is out of scope

    def render(self, *a, **k):
        return None

    def close(self):
        self.h.close()


def make(env_id="gym_kinova_gripper:kinovagripper-v0", **kw):
    if not env_id.endswith("kinovagripper-v0"):
        raise ValueError(f"unknown env id {env_id}")
    return KinovaGripper_Env(arm_or_end_effector="hand", frame_skip=15, **kw)
