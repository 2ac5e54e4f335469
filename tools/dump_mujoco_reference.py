#!/usr/bin/env python
"""Record golden tuples from the REAL reference (gym_kinova_gripper on mujoco-py 1.50.1.0 / MuJoCo 1.50) -- SURVEY 8c, "when
a real oracle exists".  Run it on a machine that has mjpro150 + licence key + mujoco-py, from the reference checkout's
gym-kinova-gripper directory (the env resolves its pose files relative to the CWD):

    cd <KinovaGrasping>/gym-kinova-gripper && python <this repo>/tools/dump_mujoco_reference.py --out <this repo>/tests/golden/mujoco150_steps.npz

It cannot run in the build container (closed binary, no network), so nothing under tests/ depends on its output being
there: tests/test_mujoco_pin.py is skipped until the file exists, and from then on pins the fp64 oracle -- and through it
the CUDA path -- to MuJoCo itself with BASELINE.json's tolerances.

Per episode (shape x orientation x seed) and RL step it stores, straight from `sim.data` / the env's own return values:
qpos/qvel before and after `env.step`, the 4-d action, ctrl after the action mapping, obs (82) before and after, reward,
done, geom_xpos / geom_xmat / site_xpos / sensordata of the last forward pass, ncon; and per episode the shape name, the
orientation, the pose-file row, the object start position and the hand euler string the env wrote into its XML."""
import argparse
import re

import numpy as np

SHAPES = ["CubeS", "CubeB", "CylinderS", "CylinderB", "Cube45S", "Cube45B", "Cone1S", "Cone1B",
          "Cone2S", "Cone2B", "Vase1S", "Vase1B", "Vase2S", "Vase2B"]


def xml_euler(env):
    text = open(env.file_dir + env.filename).read()
    at = text.find('<body name="j2s7s300_link_7"')
    m = re.search(r'euler="([^"]+)"', text[at:])
    return [float(v) for v in m.group(1).split()]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--episodes-per-shape", type=int, default=6)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    import gym
    env = gym.make("gym_kinova_gripper:kinovagripper-v0")
    raw = env.env
    rec = {k: [] for k in ("episode", "step", "qpos0", "qvel0", "qpos1", "qvel1", "action", "ctrl", "obs0", "obs1", "reward", "done",
                           "geom_xpos", "geom_xmat", "site_xpos", "sensordata", "ncon")}
    meta = {k: [] for k in ("shape", "orientation", "row", "obj_pos", "hand_euler")}
    ep = 0
    for si, shape in enumerate(SHAPES):
        for k in range(args.episodes_per_shape):
            np.random.seed(args.seed + 1000 * si + k)
            orientation = ["normal", "rotated", "top"][k % 3]
            obs = env.reset(shape_keys=[shape], hand_orientation=orientation, obj_params=[shape[:-1], shape[-1]], mode="train")
            meta["shape"].append(si); meta["orientation"].append(k % 3); meta["row"].append(int(raw.get_orientation_idx()))
            meta["obj_pos"].append(np.array(raw.get_obj_coords(), dtype=np.float64)); meta["hand_euler"].append(np.array(xml_euler(raw)))
            for t in range(args.steps):
                sim = raw._sim
                q0, v0 = sim.data.qpos.copy(), sim.data.qvel.copy()
                action = np.array([0.0, 0.0, 0.0, 0.0]) if t == 0 else np.clip(np.random.uniform(0.0, 0.8, 4) * [0.3, 1, 1, 1], 0.0, 0.8)
                nxt, reward, done, info = env.step(action)
                sim = raw._sim
                rec["episode"].append(ep); rec["step"].append(t)
                rec["qpos0"].append(q0); rec["qvel0"].append(v0); rec["qpos1"].append(sim.data.qpos.copy()); rec["qvel1"].append(sim.data.qvel.copy())
                rec["action"].append(action); rec["ctrl"].append(sim.data.ctrl.copy())
                rec["obs0"].append(np.array(obs[:82], dtype=np.float64)); rec["obs1"].append(np.array(nxt[:82], dtype=np.float64))
                rec["reward"].append(float(reward)); rec["done"].append(bool(done))
                rec["geom_xpos"].append(sim.data.geom_xpos.copy()); rec["geom_xmat"].append(sim.data.geom_xmat.copy())
                rec["site_xpos"].append(sim.data.site_xpos.copy()); rec["sensordata"].append(sim.data.sensordata.copy())
                rec["ncon"].append(int(sim.data.ncon))
                obs = nxt
                if done:
                    break
            ep += 1
    out = {k: np.array(v) for k, v in rec.items()}
    out.update({"meta_" + k: np.array(v) for k, v in meta.items()})
    np.savez_compressed(args.out, **out)
    print(f"wrote {args.out}: {ep} episodes, {len(rec['step'])} RL steps")


if __name__ == "__main__":
    main()
