"""Test-only harness that runs the REFERENCE's own rollout loops, unmodified, on this repo's env facade.

Only usable where /root/reference is mounted (the authoring container); the GPU box has no copy, so the call-site
tests that need it are CPU tests (host emulation of the kernels) and skip themselves elsewhere.

* install_stubs(make_env)  -- sys.modules stand-ins for what the reference imports but this image lacks (gym, mujoco_py,
                              glfw, matplotlib, PIL, tensorboardX, classifier_network, its matplotlib-based plotting
                              modules); gym.make -> make_env.
* import_reference(name)   -- import expert_data / main_DDPGfD / utils / DDPGfD from /root/reference, untouched.
* OracleHandle             -- the subset of kinovagrasping_b200.handle.Handle the facade uses, served by the fp64 oracle:
                              the same facade on top of it is the "oracle-driven run" the product run is compared with.
"""
import contextlib
import importlib
import io
import os
import sys
import types

import numpy as np

REF = "/root/reference/gym-kinova-gripper"


def available():
    return os.path.isdir(REF)


class _Any:
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Any()
    def __getattr__(self, n): return _Any()


def install_stubs(make_env):
    def mod(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    mod("mujoco_py", MjViewer=_Any, load_model_from_path=_Any, MjSim=_Any, MjRenderContextOffscreen=_Any)
    gym = mod("gym", Env=object, make=lambda *a, **k: make_env())

    class Box:
        def __init__(self, low, high, dtype=None, shape=None):
            self.low, self.high = np.asarray(low), np.asarray(high)
            self.shape = self.low.shape
    gym.spaces = mod("gym.spaces", Box=Box)
    gym.wrappers = mod("gym.wrappers")
    gym.utils = mod("gym.utils", EzPickle=object)
    gym.utils.seeding = mod("gym.utils.seeding", np_random=_Any)
    mod("glfw")
    mpl = mod("matplotlib", use=lambda *a, **k: None)
    mpl.pyplot = mod("matplotlib.pyplot")
    mod("classifier_network", LinearNetwork=_Any, ReducedLinearNetwork=_Any)
    mod("PIL", Image=_Any(), ImageFont=_Any(), ImageDraw=_Any())
    mod("PIL.Image"); mod("PIL.ImageFont"); mod("PIL.ImageDraw")
    mod("tensorboardX", SummaryWriter=_Any)
    # matplotlib-based plotting modules; heatmap_coords is plain numpy and is imported for real
    for n in ("heatmap_plot", "boxplot_plot", "trajectory_plot"):
        mod(n, generate_heatmaps=_Any, generate_reward_boxplots=_Any, plot_trajectory=_Any)
    for p in (REF, os.path.join(REF, "plotting_code")):
        if p not in sys.path:
            sys.path.insert(0, p)


def import_reference(name):
    """Import (or re-import, so that module-level state is fresh) one of the reference's modules by bare name."""
    if name in sys.modules and getattr(sys.modules[name], "__file__", "").startswith(REF):
        return importlib.reload(sys.modules[name])
    return importlib.import_module(name)


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


class OracleHandle:
    """kinovagrasping_b200.handle.Handle as the facade sees it, on the fp64 oracle (one env)."""

    def __init__(self):
        from oracle import kgo
        self.kgo, self.m = kgo, kgo.Model()
        self.n = 1
        self.e = None
        self.frame_skip = 15
        self._obs = np.zeros(82)

    def set_frame_skip(self, n): assert n == 15
    def set_action_dim(self, d): assert d == 6
    def set_max_episode_steps(self, n): pass
    def enable_obs_global(self): pass
    def close(self): pass

    def reset_host(self, shape, euler, objpos, mask=None, handpos=None):
        hp = None if handpos is None else np.asarray(handpos, dtype=np.float64).reshape(3)
        self.e, obs = self.m.env_reset(int(np.asarray(shape).reshape(-1)[0]), np.asarray(euler, dtype=np.float64).reshape(3),
                                       np.asarray(objpos, dtype=np.float64).reshape(3), hand_pos=hp)
        self.e.max_steps = 10 ** 9
        self._obs = obs
        return obs.astype(np.float32).reshape(82, 1)

    def step_host(self, action=None, expert_mode=0, **kw):
        obs, r, d, _ = self.m.env_step(self.e, np.asarray(action, dtype=np.float64).reshape(6))
        self._obs = obs
        return obs.astype(np.float32).reshape(82, 1), np.array([r], dtype=np.float32), np.array([int(d)], dtype=np.int32)

    def get_state_host(self):
        return (np.array(self.e.st.qpos, dtype=np.float32).reshape(16, 1), np.array(self.e.st.qvel, dtype=np.float32).reshape(15, 1))

    def obs_global_host(self):
        return self.m.get_obs(self.e.st, self.e.d, True).astype(np.float32).reshape(74, 1)
