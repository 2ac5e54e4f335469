"""Batched evaluation with the reference's eval outputs (main_DDPGfD.py:130-272 eval_policy and what
update_policy does with its result, MAIN:486-528).

`eval_policy` runs ONE lock-step episode of every env of a VecKinovaGripperEnv -- the policy acts until check_grasp's
finger-distance test hands over to the scripted lift (MAIN:186-229; the same hand-over as train.run_episode) -- and returns
the dictionary MAIN:271 returns, one "evaluation episode" per env:
    avg_reward, avg_rewards{total,finger,grasp,lift}, all_ep_reward_values{...}[E], num_success, success_coords / fail_coords
    {"x", "y", "orientation"} = the object's start position in the hand frame (MAIN:171-174: Tfw @ [obj_coords, 1], which is
    obs[21:24] of the reset observation) split by lift success (reward > 25, MAIN:236).
Reward bookkeeping follows the reference's quirk: during the scripted lift a step's reward is only counted on the step
that ends the episode (MAIN:222-229).

`save_eval_outputs` writes the files the reference's plotting code reads: heatmap coordinates per orientation
(plotting_code/heatmap_coords.py:33-95 -- <dir>/<orientation>/{success,fail,total}_{x,y}_<episode>.npy and heatmap_info.txt)
and the four boxplot arrays (MAIN:525-528, the `[[values]]` nesting of MAIN:361-364, 510-513 included)."""
import os

import numpy as np

ORIENTATION_NAMES = ("normal", "rotated", "top")


def eval_policy(env, act_fn, shapes=("CubeS",), orientation="normal", variant="A", mode="test", seed=0, max_num_timesteps=30):
    import torch as t
    from .train import FINGER_LIFT_VELOCITY, SKIP_NUM_TS, WRIST_LIFT_VELOCITY, settled
    obs = env.reset(shapes, orientation, mode=mode, variant=variant, seed=seed).clone()
    n, dev = env.n, env.device
    start_local = obs[:, 21:24].clone()                      # Tfw @ [obj_coords, 1]  (MAIN:171-174)
    orient = env.h.orientation_host()
    alive = t.ones(n, dtype=t.bool, device=dev)
    ready = t.zeros(n, dtype=t.bool, device=dev)
    success = t.zeros(n, dtype=t.bool, device=dev)
    sums = {k: t.zeros(n, device=dev) for k in ("total_reward", "finger_reward", "grasp_reward", "lift_reward")}
    lift = t.tensor([WRIST_LIFT_VELOCITY, FINGER_LIFT_VELOCITY, FINGER_LIFT_VELOCITY, FINGER_LIFT_VELOCITY], device=dev)
    prev = None
    steps = min(int(max_num_timesteps), int(env._max_episode_steps))
    for step in range(1, steps + 1):
        if prev is not None and step >= SKIP_NUM_TS:
            ready |= settled(prev, obs)
        action = t.where(ready[:, None], lift[None, :], act_fn(obs))
        nxt, rew, done, info = env.step(action)
        if step == steps:
            done = t.ones_like(done)                          # eval_env._max_episode_steps = max_num_timesteps (MAIN:167)
        counted = alive & (~ready | done)                     # policy steps always; lift steps only when they end the episode
        for key, val in (("total_reward", rew), ("finger_reward", info["finger_reward"]), ("grasp_reward", info["grasp_reward"]),
                         ("lift_reward", info["lift_reward"])):
            v = val if t.is_tensor(val) else t.full((n,), float(val), device=dev)
            sums[key] += t.where(counted, v, t.zeros_like(v))
        success |= alive & (rew > 25)
        alive &= ~done
        prev, obs = obs, nxt.clone()
    ok = success.cpu().numpy()
    xy = start_local.cpu().numpy().astype(np.float64)
    names = np.array([ORIENTATION_NAMES[o] if 0 <= o < 3 else str(orientation) for o in orient])
    vals = {k: v.cpu().numpy().astype(np.float64) for k, v in sums.items()}
    return {"avg_reward": float(vals["total_reward"].mean()),
            "avg_rewards": {k: float(np.average(v)) for k, v in vals.items()},
            "all_ep_reward_values": {k: list(v) for k, v in vals.items()},
            "num_success": int(ok.sum()),
            "success_coords": {"x": xy[ok, 0], "y": xy[ok, 1], "orientation": names[ok]},
            "fail_coords": {"x": xy[~ok, 0], "y": xy[~ok, 1], "orientation": names[~ok]}}


def save_eval_outputs(ret, episode_num, heatmap_dir, boxplot_dir):
    """The files MAIN:516-528 leaves behind for one evaluation (read by heatmap_plot.py / boxplot_plot.py)."""
    os.makedirs(heatmap_dir, exist_ok=True)
    os.makedirs(boxplot_dir, exist_ok=True)
    ep = "" if episode_num is None else "_" + str(episode_num)
    counts = {}
    for kind, coords in (("success", ret["success_coords"]), ("fail", ret["fail_coords"])):
        ori = np.asarray(coords["orientation"])
        for name in ORIENTATION_NAMES:
            idx = np.flatnonzero(ori == name)
            counts[(kind, name)] = len(idx)
            if len(idx) == 0:
                continue
            d = os.path.join(heatmap_dir, name)
            os.makedirs(d, exist_ok=True)
            x, y = np.asarray(coords["x"])[idx], np.asarray(coords["y"])[idx]
            # coords_dict_to_array is called for "/success" (or "/fail") and then for "/total" with the SAME index list, so the
            # "total" files of an orientation end up holding whichever of the two was written last: the failures if there are any
            np.save(os.path.join(d, kind + "_x" + ep), x); np.save(os.path.join(d, kind + "_y" + ep), y)
            np.save(os.path.join(d, "total_x" + ep), x); np.save(os.path.join(d, "total_y" + ep), y)
    with open(os.path.join(heatmap_dir, "heatmap_info.txt"), "w") as f:
        f.write("Heatmap Coords \nSaved at: " + heatmap_dir + "\n\nTotal # Success: " + str(len(ret["success_coords"]["x"])) +
                "\nTotal # Fail: " + str(len(ret["fail_coords"]["x"])) + "\n")
        for name, title in zip(ORIENTATION_NAMES, ("Normal", "Rotated", "Top")):
            f.write("\n" + title + " Orientation\n# Success: " + str(counts[("success", name)]) + "\n# Fail: " + str(counts[("fail", name)]) + "\n")
    for key in ("finger_reward", "grasp_reward", "lift_reward", "total_reward"):
        np.save(os.path.join(boxplot_dir, key + ep), [[ret["all_ep_reward_values"][key]]])
