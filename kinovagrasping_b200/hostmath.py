"""Small host-side (numpy) helpers for attributes callers read off the env object (env.Tfw, env.wrist_pose)."""
import numpy as np


def _q2m(q):
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def euler2mat(e):
    """MuJoCo eulerseq 'xyz' (intrinsic)."""
    cx, sx, cy, sy, cz, sz = np.cos(e[0]), np.sin(e[0]), np.cos(e[1]), np.sin(e[1]), np.cos(e[2]), np.sin(e[2])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rx @ Ry @ Rz


def hand_frame_from_euler(card, euler, slides):
    """Tfw (4x4) and wrist_pose as _get_trans_mat_wrist_pose computes them (kinova_gripper_env.py:274-288)."""
    Rh = euler2mat(np.asarray(euler, dtype=np.float64))
    axes = card.get("slide_axes").reshape(3, 3)
    hp = card.get("l7_pos") + Rh @ (axes.T @ np.asarray(slides, dtype=np.float64))
    mg = card.get("mesh_geom").reshape(-1, 15)[0]
    palm_pos = hp + Rh @ mg[0:3]
    Rfa = Rh @ _q2m(mg[3:7])
    temp = (Rfa @ np.array([[0, 0, 1], [-1, 0, 0], [0, -1, 0]])).T
    Tfw = np.zeros((4, 4))
    Tfw[0:3, 0:3] = temp
    Tfw[3, 3] = 1
    wrist = palm_pos + temp.T @ np.array([-0.009, 0.048, 0.0])
    Tfw[0:3, 3] = -temp @ wrist
    return Tfw, wrist
