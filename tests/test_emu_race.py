"""Race check of the warp-phase kernels (the reference has no race detection; here the phase discipline of
kg_core.cuh is checked instead).  The host emulation runs the lanes of every phase in ascending and then in
descending order: a phase body that reads what another lane writes inside the same phase (a data race on the GPU,
where __syncwarp() only separates phases) produces different results.  Outputs must be bit-identical."""
import numpy as np

from kinovagrasping_b200.handle import Handle


def _rollout(lib, card, order):
    lib.kg_emu_set_lane_order(order)
    n = 14
    pose = card.pose("train")
    shape = np.arange(n, dtype=np.int32) % 14
    rows = 17 + 31 * np.arange(n)
    objp = np.stack([pose[shape[i], 0, rows[i], 0:3] for i in range(n)], 1).astype(np.float32)
    euler = np.stack([pose[shape[i], 2, rows[i], 6:9] for i in range(n)], 1).astype(np.float32)
    h = Handle(n, lib=lib)
    out = [h.reset_host(shape, euler, objp).copy()]
    for t in range(26):
        obs, rew, done = h.step_host(None, 2)
        out.append(obs.copy()); out.append(rew.copy()); out.append(done.astype(np.float32))
    qp, qv = h.get_state_host()
    lib.kg_emu_set_lane_order(0)
    return out + [qp, qv]


def test_phase_discipline_lane_order_invariance(card):
    from conftest import _emu_lib
    lib = _emu_lib()
    a = _rollout(lib, card, 0)
    b = _rollout(lib, card, 1)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
