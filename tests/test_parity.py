"""Parity of the warp-per-env kernels against the fp64 oracle, through the C ABI (include/kg.h).
Every test runs twice: backend 'emu' (host emulation of the same kernel source; CPU suite) and backend 'cuda'
(@gpu: the product on the B200).  Tolerances follow BASELINE.json's north_star:
  obs 1e-5, reward/done exact, one-step qpos/qvel 1e-4 relative (fp32 vs fp64) where the narrowphase is
  well-conditioned; MPR contact normals at inflated edges are only defined to a few degrees at mpr_tolerance=1e-6
  (DESIGN.md "Conditioning of the narrowphase"), so contact-rich states are checked stage by stage instead."""
import ctypes as C

import numpy as np
import pytest

from kinovagrasping_b200.handle import Handle
from oracle import kgo

SIDE = np.array([-1.57, 0.0, -1.57])


def _poses(card, n, seed, orient_idx=(0, 1, 2), shapes=range(14)):
    rng = np.random.default_rng(seed)
    pose = card.pose("train")
    shape = np.array([list(shapes)[i % len(list(shapes))] for i in range(n)], dtype=np.int32)
    rows = rng.integers(1, 4500, n)
    ori = np.array([orient_idx[i % len(orient_idx)] for i in range(n)])
    objp = np.stack([pose[shape[i], ori[i], rows[i], 0:3] for i in range(n)], 1).astype(np.float32)
    euler = np.stack([pose[shape[i], ori[i], rows[i], 6:9] for i in range(n)], 1).astype(np.float32)
    return shape, euler, objp


def _oracle_reset(oracle, shape, euler, objp):
    out = []
    for i in range(len(shape)):
        out.append(list(oracle.env_reset(int(shape[i]), euler[:, i].astype(np.float64), objp[:, i].astype(np.float64))))
    return out


def _obs_err(a, b):
    return np.abs(a - b) / np.maximum(1.0, np.abs(b))


def test_reset_obs_parity_all_shapes_orientations(backend, oracle, card):
    n = 144                                         # 24 shapes (the 14 of Exp1, the 7 medium sizes, the 3 lemons) x 3 orientations x 2 rows
    shape, euler, objp = _poses(card, n, 1, shapes=range(24))
    h = Handle(n, lib=backend)
    obs = h.reset_host(shape, euler, objp)
    ref = _oracle_reset(oracle, shape, euler, objp)
    worst = max(_obs_err(obs[:, i].astype(np.float64), ref[i][1]).max() for i in range(n))
    assert worst < 1e-5, worst
    assert np.abs(obs[18:21]).max() < 1e-6          # wrist in its own frame
    qpos, qvel = h.get_state_host()
    assert np.all(qvel == 0) and np.all(qpos[12] == 1)


def test_smooth_dynamics_parity(backend, oracle, card):
    """M (CRBA), qfrc_bias (RNE) and qacc_smooth from random states: closed-form kernels vs generic spatial algebra."""
    n = 28
    shape, euler, objp = _poses(card, n, 2, orient_idx=(1, 2))
    objp[2] += 0.3                                   # lift the objects clear of the hand and the floor
    h = Handle(n, lib=backend)
    h.reset_host(shape, euler, objp)
    rng = np.random.default_rng(3)
    qpos, qvel = h.get_state_host()
    qpos[0:3] = rng.uniform(-0.05, 0.05, (3, n))
    prox = rng.uniform(0, 1.2, (3, n))
    qpos[3:9:2] = prox; qpos[4:9:2] = prox / 2 + rng.normal(0, 0.02, (3, n))
    quat = np.array([1, 0, 0, 0])[:, None] + rng.normal(0, 0.3, (4, n)); quat /= np.linalg.norm(quat, axis=0)
    qpos[12:16] = quat
    qvel = rng.normal(0, 0.5, (15, n)).astype(np.float32)
    qvel[12:15] *= 4
    h.set_state_host(qpos, qvel)
    ctrl = np.array([0.1, 0.2, -0.1, 0.1, 0.05, 0.29, 0.5, 0.4, 0.3])
    nfree = 0
    for i in range(n):
        d = h.forward_debug(i, ctrl)
        s = oracle.state(int(shape[i]), euler[:, i].astype(np.float64), objp[:, i].astype(np.float64))
        for k in range(16): s.qpos[k] = float(qpos[k, i])
        for k in range(15): s.qvel[k] = float(qvel[k, i])
        D = oracle.forward(s, ctrl)
        M = D.arr("M", (15, 15))
        assert D.ncon == d["ncon"]
        assert np.abs(M - d["M"]).max() < 1e-6 * np.abs(M).max()
        b = D.arr("qfrc_bias")
        assert np.abs(b - d["bias"]).max() < 2e-5 * max(1.0, np.abs(b).max())
        qs = D.arr("qacc_smooth")
        assert np.abs(qs - d["qacc_smooth"]).max() < 1e-4 * max(1.0, np.abs(qs).max())
        if D.ncon == 0:                               # 3 tendon rows are always present -> exercises the Newton solver
            qa = D.arr("qacc"); nfree += 1
            assert np.abs(qa - d["qacc"]).max() < 1e-4 * max(1.0, np.abs(qa).max()), (i, np.abs(qa - d["qacc"]).max())
    assert nfree >= 5


_LOCKSTEP_CACHE = {}


def _lockstep(backend, oracle, card, n, seed, steps, mode=0, shapes=(0, 1)):
    key = (backend._name, n, seed, steps, mode, tuple(shapes))
    if key not in _LOCKSTEP_CACHE:
        _LOCKSTEP_CACHE[key] = _lockstep_run(backend, oracle, card, n, seed, steps, mode, shapes)
    return _LOCKSTEP_CACHE[key]


def _lockstep_run(backend, oracle, card, n, seed, steps, mode=0, shapes=(0, 1)):
    """Oracle-driven side-grasp rollouts.  Before EVERY physics substep the kernel state is overwritten with the
    oracle's (fp32-rounded) state, so each record is a true one-step (one mj_step) error from identical
    qpos/qvel/ctrl, which is what the north_star tolerance is stated on.  With frame_skip = 1 every call also runs the
    obs epilogue (rangefinder BVH, _get_obs local + global, experimental_sensor) on the kinematics of that same state, so
    each record also carries the post-step observation error of all 82 / 74 entries against the oracle's _get_obs."""
    shape, _, objp = _poses(card, n, seed, orient_idx=(0,), shapes=shapes)
    _, euler, _ = _poses(card, n, seed, orient_idx=(2,), shapes=shapes)
    h = Handle(n, lib=backend)
    h.set_max_episode_steps(10 ** 6)
    h.set_frame_skip(1)
    h.enable_obs_global()
    h.reset_host(shape, euler, objp)
    ref = _oracle_reset(oracle, shape, euler, objp)
    dp = C.POINTER(C.c_double)
    recs = []
    for t in range(steps):
        acts = np.zeros((4, n), dtype=np.float32); ctrls = []
        for i in range(n):
            a, _ = oracle.expert_action(ref[i][0], ref[i][1], mode)
            acts[:, i] = a
            c = np.zeros(9)
            kgo.lib().kgo_ctrl_from_action(C.byref(ref[i][0].d), a.ctypes.data_as(dp), c.ctypes.data_as(dp))
            ctrls.append(c)
        for sub in range(15):
            qp = np.stack([np.array(ref[i][0].st.qpos) for i in range(n)], 1).astype(np.float32)
            qv = np.stack([np.array(ref[i][0].st.qvel) for i in range(n)], 1).astype(np.float32)
            h.set_state_host(qp, qv)
            dbg = [h.forward_debug(i, ctrls[i]) for i in range(n)]      # the kernel's contact set for this state
            obs, rew, done = h.step_host(acts, 0)
            obsg = h.obs_global_host()
            qp2, qv2 = h.get_state_host()
            for i in range(n):
                e = ref[i][0]
                for k in range(16): e.st.qpos[k] = float(qp[k, i])
                for k in range(15): e.st.qvel[k] = float(qv[k, i])
                st0 = kgo.State.from_buffer_copy(e.st)
                D = oracle.step(e.st, ctrls[i], rays=True)
                e.d = D
                # the same step once more with the contact list the KERNEL found for this state: what is left is P6-P9 alone
                evg = eqg = None
                if dbg[i]["ncon"] <= 23:
                    given = []
                    for c in dbg[i]["con"]:
                        code = int(round(float(c[7]) * 10))
                        given.append([(code // 100) % 10, code // 1000, float(c[0])] + [float(x) for x in c[1:7]])
                    sg = kgo.State.from_buffer_copy(st0)
                    oracle.step_contacts(sg, ctrls[i], np.array(given).reshape(-1, 9))
                    evg = np.abs(qv2[:, i] - np.array(sg.qvel)).max(); eqg = np.abs(qp2[:, i] - np.array(sg.qpos)).max()
                # _get_obs reads the kinematics / sensors of the forward pass (pre-integration) and data.qpos-independent
                # quantities only, so the state handed to it is the one the forward pass ran on
                o_l = oracle.get_obs(st0, D, False); o_g = oracle.get_obs(st0, D, True)
                oerr = _obs_err(obs[:, i].astype(np.float64), o_l); gerr = _obs_err(obsg[:, i].astype(np.float64), o_g)
                same = (D.ncon == dbg[i]["ncon"])
                if same:
                    for c in range(D.ncon):
                        cc = D.con[c]
                        ok = any(np.abs(k[1:4] - np.array(cc.pos)).max() < 2e-5 and abs(float(k[0]) - cc.dist) < 2e-6 and
                                 float(np.dot(k[4:7], np.array(cc.frame)[:3])) > 0.99996 for k in dbg[i]["con"])
                        same = same and ok
                z = D.geom_xpos[26]
                r = 50.0 if (abs(z - 0.2) < 0.005 or z >= 0.2) else 0.0
                recs.append(dict(t=t, i=i, ncon=D.ncon, hull=sum(1 for c in range(D.ncon) if D.con[c].geom1 != 0),
                                 eq=np.abs(qp2[:, i] - np.array(e.st.qpos)).max(),
                                 ev=np.abs(qv2[:, i] - np.array(e.st.qvel)).max(), vs=np.abs(np.array(e.st.qvel)).max(),
                                 rew_ok=(float(rew[i]) == r), z=z, same=same, oerr=oerr, gerr=gerr, obs_ref=o_l, obs=obs[:, i].copy(), evg=evg, eqg=eqg,
                                 mind=min([D.con[c].dist for c in range(D.ncon)] + [1.0])))
        for i in range(n):
            ref[i][1] = oracle.get_obs(ref[i][0].st, ref[i][0].d, False)
    return recs


def test_one_step_parity_same_contact_set(backend, oracle, card):
    """north_star gate: one mj_step from identical qpos/qvel/ctrl, next qpos/qvel within 1e-4 relative (fp32 kernel vs
    fp64 oracle; scale max(1, |qvel|_inf)) whenever both narrowphases report the same contact set (incl. none).
    Covers free flight, resting/landing plane contacts, the tendon equality rows and finger-object contacts."""
    recs = _lockstep(backend, oracle, card, n=10, seed=5, steps=24)
    # deep initial interpenetration (some variant-B table rows start with the object inside a finger) is a stiff,
    # non-physical transient where MPR's answer is path dependent: excluded from the strict gate, kept in the statistics test
    same = [r for r in recs if r["same"] and r["mind"] > -0.002]
    with_hull = [r for r in same if r["hull"] > 0]
    print(f"substeps {len(recs)}, same contact set {len(same)} ({len(with_hull)} of them with finger/palm-object contacts)")
    assert len(same) > 0.6 * len(recs) and len(with_hull) > 100
    ev = np.array([r["ev"] / max(1.0, r["vs"]) for r in same])
    nocon = [r for r in same if r["ncon"] == 0]
    print(f"same-set substeps within 1e-4: {np.mean(ev < 1e-4):.4f}, max {ev.max():.2e}; contact-free substeps {len(nocon)}")
    eq = np.array([r["eq"] for r in same])
    # "same set" (judged on a separately compiled dump kernel, to 0.5 degrees / 2e-6 m) still leaves room for the two
    # narrowphases to differ slightly; a firm grip turns that into a few 1e-4 of velocity.  Gate: >= 95 % of these
    # substeps inside 1e-4, >= 99 % inside 1e-3, and EVERY contact-free substep inside 1e-4.
    assert np.mean((ev < 1e-4) & (eq < 1e-4)) >= 0.95, np.mean((ev < 1e-4) & (eq < 1e-4))
    assert np.mean((ev < 1e-3) & (eq < 1e-4)) >= 0.99, np.mean((ev < 1e-3) & (eq < 1e-4))
    assert all(r["rew_ok"] or min(abs(r["z"] - 0.195), abs(r["z"] - 0.2)) < 1e-4 for r in same)
    assert len(nocon) > 50
    for r in nocon:
        assert r["eq"] < 1e-4 and r["ev"] < 1e-4 * max(1.0, r["vs"]), r


def test_one_step_parity_every_substep_on_identical_contacts(backend, oracle, card):
    """north_star gate without the narrowphase's conditioning: one mj_step from identical qpos/qvel/ctrl AND the contact list
    the kernel itself found for that state (oracle: kgo_step_contacts) -- next qpos/qvel within 1e-4 (qvel relative to
    max(1, |qvel|_inf)) on EVERY substep of the lock-step rollouts: free flight, landing, tendon rows, limits, grip, lift."""
    recs = _lockstep(backend, oracle, card, n=10, seed=5, steps=24) + _lockstep(backend, oracle, card, n=8, seed=7, steps=24)
    got = [r for r in recs if r["evg"] is not None]
    assert len(got) > 0.99 * len(recs) and len(got) > 6000
    ev = np.array([r["evg"] / max(1.0, r["vs"]) for r in got]); eq = np.array([r["eqg"] for r in got])
    ncon = np.array([r["ncon"] for r in got])
    print(f"identical contacts: {len(got)} substeps ({int((ncon > 0).sum())} with contacts, up to {ncon.max()}): qvel err max {ev.max():.2e}, "
          f"p99 {np.percentile(ev, 99):.2e}; qpos err max {eq.max():.2e}")
    assert ev.max() < 1e-4 and eq.max() < 1e-4, (ev.max(), eq.max())


def test_medium_size_shapes_lockstep(backend, oracle, card):
    """SURVEY 8(f4) 'remaining shapes': the seven medium sizes (CubeM ... Vase2M, ENV:150-178; MAIN:1271 tests on them) and the
    three lemons (ENV:198-201: a 2434-vertex hull, the geom rotated in the body, the `'lemon' in self.filename` size rule) through
    the same two gates as the Exp1 shapes -- one mj_step on identical contacts within 1e-4 on every substep, and the 60 smooth
    post-step obs entries within 1e-5 on every substep (rangefinders >= 99.5 %)."""
    recs = _lockstep(backend, oracle, card, n=10, seed=11, steps=20, shapes=tuple(range(14, 24)))
    got = [r for r in recs if r["evg"] is not None]
    assert len(got) > 0.99 * len(recs)
    ev = np.array([r["evg"] / max(1.0, r["vs"]) for r in got]); eq = np.array([r["eqg"] for r in got])
    ncon = np.array([r["ncon"] for r in got])
    oerr = np.stack([r["oerr"] for r in recs])
    smooth = [k for k in range(82) if k not in list(range(50, 67)) + list(range(70, 75))]
    print(f"medium shapes: {len(got)} substeps ({int((ncon > 0).sum())} with contacts, up to {ncon.max()}): qvel err max {ev.max():.2e}, "
          f"qpos {eq.max():.2e}, smooth obs {oerr[:, smooth].max():.2e}, rays within 1e-5 {np.mean(oerr[:, 50:67] < 1e-5):.4f}")
    assert (ncon > 0).sum() > 1500
    # The pose tables put a lemon's BODY origin where the other shapes have their centre, but its geom sits 5-7 cm off the body
    # origin (Lemon_Standin.STL is not centred), so a lemon episode starts with the fruit a few mm inside the floor and the first
    # substeps throw it out at ~0.85 m/s: stiff de-penetration, where fp32 is held to 3e-4; every other substep to 1e-4.
    deep = np.array([r["mind"] < -0.002 for r in got])
    print(f"   substeps with a contact deeper than 2 mm: {int(deep.sum())}, qvel err max there {ev[deep].max() if deep.any() else 0:.2e}; elsewhere {ev[~deep].max():.2e}")
    assert ev[~deep].max() < 1e-4 and eq.max() < 1e-4, (ev[~deep].max(), eq.max())
    assert deep.sum() < 0.02 * len(got) and (not deep.any() or ev[deep].max() < 3e-4)
    assert oerr[:, smooth].max() < 1e-5, oerr[:, smooth].max()
    assert (oerr[:, 50:67] < 1e-5).mean() >= 0.995


RAY_ENTRIES = list(range(50, 67)) + list(range(70, 75))     # rangefinders and what experimental_sensor derives from them


def test_post_step_obs_parity_all_82_and_74(backend, oracle, card):
    """north_star gate "observation vectors within 1e-5": the obs the kernel emits after EVERY lock-step substep (same
    state, same ctrl as the oracle; fingers closing on the object, contacts, lift) against the oracle's _get_obs -- all
    82 local and all 74 global entries, tolerance 1e-5 relative to max(1, |x|).
    The 60 entries that are smooth functions of the kinematics must hold on every substep.  The 17 rangefinders (and the
    5 entries experimental_sensor derives from them) are piecewise: a ray that grazes a triangle edge or a silhouette
    can land on a different triangle in fp32 and fp64, so they are held to 1e-5 on >= 99.5 % of the (substep, entry)
    pairs, and every miss must be such a discontinuity (the ray's own range jumps by > 1 mm between neighbouring
    substeps in the ORACLE's own trace) -- never a small smooth error."""
    recs = _lockstep(backend, oracle, card, n=10, seed=5, steps=24)
    oerr = np.stack([r["oerr"] for r in recs]); gerr = np.stack([r["gerr"] for r in recs])
    smooth = [k for k in range(82) if k not in RAY_ENTRIES]
    worst = int(np.argmax(oerr[:, smooth].max(axis=0)))
    print(f"post-step obs, {len(recs)} substeps: smooth entries max rel err {oerr[:, smooth].max():.2e} (entry {smooth[worst]}); "
          f"ray-derived entries within 1e-5: {np.mean(oerr[:, RAY_ENTRIES] < 1e-5):.5f}, max {oerr[:, RAY_ENTRIES].max():.2e}")
    assert oerr[:, smooth].max() < 1e-5, (smooth[worst], oerr[:, smooth].max())
    gsmooth = [k for k in range(74) if not (50 <= k < 67)]
    assert gerr[:, gsmooth].max() < 1e-5, gerr[:, gsmooth].max()
    ray_ok = oerr[:, 50:67] < 1e-5
    assert ray_ok.mean() >= 0.995, ray_ok.mean()
    assert np.array_equal(gerr[:, 50:67] < 1e-5, ray_ok)               # the global layout carries the same 17 numbers
    # rays pointing at closed fingers / the object in the hand really are exercised (only reset was compared)
    rng_ref = np.stack([r["obs_ref"][50:67] for r in recs])
    assert (rng_ref < 0.06).mean() > 0.05 and (rng_ref[:, 5:] < 6.0).mean() > 0.2
    # every miss is a hit/miss flip or a jump between surfaces, not a smooth-error excursion
    bad = np.argwhere(~ray_ok)
    for k, j in bad:
        a, b = recs[k]["obs"][50 + j], recs[k]["obs_ref"][50 + j]
        assert abs(a - b) > 2e-4, (k, j, a, b)
    # experimental_sensor branches after motion: range < 0.06 taken for some palm rays, and the derived entries agree
    # wherever the five palm rays agree
    palm_ok = ray_ok[:, :5].all(axis=1)
    assert (rng_ref[:, :5] < 0.06).any(axis=1).mean() > 0.05
    assert oerr[palm_ok][:, 70:75].max() < 1e-5, oerr[palm_ok][:, 70:75].max()


def test_one_step_parity_all_substeps_statistics(backend, oracle, card):
    """Every substep, including those where tie-breaks between equally deep hull vertices or the few-degree freedom of an
    MPR normal at an inflated edge (DESIGN.md "Conditioning of the narrowphase") give the two narrowphases different
    contact sets: the bulk stays inside 1e-4, the tail is bounded."""
    recs = _lockstep(backend, oracle, card, n=8, seed=7, steps=24)
    ev = np.array([r["ev"] / max(1.0, r["vs"]) for r in recs]); eq = np.array([r["eq"] for r in recs])
    frac = float(np.mean((ev < 1e-4) & (eq < 1e-4)))
    print(f"substeps {len(recs)}: within 1e-4: {frac:.3f}; qvel err median {np.median(ev):.2e} p99 {np.percentile(ev, 99):.2e} max {ev.max():.2e}; qpos err max {eq.max():.2e}")
    assert frac > 0.7, frac
    assert np.percentile(ev, 99) < 5e-2 and eq.max() < 5e-3, (np.percentile(ev, 99), eq.max())   # tail: differing contact sets only
    for r in recs:     # reward / done are discrete: equal unless the height sits within the one-step error of a threshold
        assert r["rew_ok"] or min(abs(r["z"] - 0.195), abs(r["z"] - 0.2)) < 1e-4, r


def test_narrowphase_against_oracle(backend, oracle, card):
    """Contact sets from identical states: same pairs, distance within 5e-6 m, position within 1e-4 m along the normal; normals:
    median < 0.05 degrees and >= 98 % within 0.1 degrees, every one within the curvature-limited cone of 8 degrees."""
    n = 8
    shape, _, objp = _poses(card, n, 11, orient_idx=(0,), shapes=(0, 1, 2, 6))
    _, euler, _ = _poses(card, n, 11, orient_idx=(2,), shapes=(0, 1, 2, 6))
    h = Handle(n, lib=backend)
    h.set_max_episode_steps(10 ** 6)
    h.reset_host(shape, euler, objp)
    ref = _oracle_reset(oracle, shape, euler, objp)
    checked = 0; ang_worst = 0.0; angs = []
    for t in range(24):
        for i in range(n):
            a, _ = oracle.expert_action(ref[i][0], ref[i][1], 0)
            ref[i][1], _, _, _ = oracle.env_step(ref[i][0], a)
        qp = np.stack([np.array(ref[i][0].st.qpos) for i in range(n)], 1).astype(np.float32)
        qv = np.stack([np.array(ref[i][0].st.qvel) for i in range(n)], 1).astype(np.float32)
        h.set_state_host(qp, qv)
        for i in range(n):
            e = ref[i][0]
            s = kgo.State.from_buffer_copy(e.st)
            for k in range(16): s.qpos[k] = float(qp[k, i])
            for k in range(15): s.qvel[k] = float(qv[k, i])
            D = oracle.forward(s, np.zeros(9))
            d = h.forward_debug(i, np.zeros(9))
            mine = {}
            for c in d["con"]:
                code = int(round(float(c[7]) * 10)); g2 = code // 1000; g1 = (code // 100) % 10
                mine.setdefault((g1, g2), []).append(c)
            theirs = {}
            for c in range(D.ncon):
                cc = D.con[c]
                theirs.setdefault((cc.geom1, cc.geom2), []).append(cc)
            for key, lst in theirs.items():
                if key[0] == 0:
                    continue          # plane contacts: ties between equally deep hull vertices are broken differently
                if abs(lst[0].dist - 0.001) < 5e-6:
                    continue          # borderline activation
                assert key in mine, (t, i, key, lst[0].dist)
                c = mine[key][0]
                assert abs(float(c[0]) - lst[0].dist) < 5e-6, (key, c[0], lst[0].dist)
                # face/edge contacts have a non-unique witness point along the contact patch: only the offset ALONG the
                # normal is pinned tightly, the lateral offset must stay inside the patch (object size)
                dpos = c[1:4] - np.array(lst[0].pos)
                assert abs(float(np.dot(dpos, np.array(lst[0].frame)[:3]))) < 1e-4 and np.abs(dpos).max() < 0.06
                cosang = float(np.dot(c[4:7], np.array(lst[0].frame)[:3]))
                ang_worst = max(ang_worst, np.degrees(np.arccos(min(1.0, cosang))))
                angs.append(np.degrees(np.arccos(min(1.0, cosang))))
                checked += 1
    assert checked > 100
    assert ang_worst < 8.0, ang_worst
    angs = np.array(angs)
    print(f"narrowphase normals, {checked} hull contacts: median {np.median(angs):.4f} deg, p90 {np.percentile(angs, 90):.4f}, p99 {np.percentile(angs, 99):.3f}, max {angs.max():.3f}")
    # the bulk is tight; the 8-degree bound above is for the rare contact ON an inflated edge, where the normal is free within
    # the edge's cone and fp32 / fp64 portal refinement stop at different points of it (conditioning)
    assert np.median(angs) < 0.05 and np.mean(angs < 0.1) >= 0.98, (np.median(angs), np.mean(angs < 0.1))


def test_constraint_rows_and_solver_on_identical_contacts(backend, oracle, card):
    """P6 and P7 without the narrowphase in between: along an expert rollout the oracle is handed the contact list THE KERNEL
    found (kgo_forward_contacts: geoms, distance, point, normal) and builds its own rows from it.  Then
      * every scalar row the solver sees -- efc_J, efc_aref, efc_D of the reference engine; kg_make_rows here, a contact's four
        pyramid rows reconstructed from its base rows as n +- mu t1, n +- mu t2 -- agrees: J 2e-6 of the row's scale, D 1e-5
        relative, aref within 1e-5 max(1, |aref|) + 16 ulp of B sum_j |J_j v_j| (aref = -B J.v - K d r is a 15-term fp32 dot
        product scaled by B = 105, so its round-off follows the magnitude of its terms; tendon equality, joint limit and
        contact rows alike);
      * the Newton solver's answer on those rows agrees: qacc within 1e-4 of max(1, |qacc|) on EVERY state, with or
        without contacts -- the north_star's one-step gate, here free of the narrowphase's conditioning."""
    n = 8
    shape, _, objp = _poses(card, n, 11, orient_idx=(0,), shapes=(0, 1, 2, 6))
    _, euler, _ = _poses(card, n, 11, orient_idx=(2,), shapes=(0, 1, 2, 6))
    h = Handle(n, lib=backend)
    h.set_max_episode_steps(10 ** 6)
    h.reset_host(shape, euler, objp)
    ref = _oracle_reset(oracle, shape, euler, objp)
    stats = {1: [], 2: [], 3: []}
    qerr = []; ferr = []; nfree = 0
    for t in range(24):
        for i in range(n):
            a, _ = oracle.expert_action(ref[i][0], ref[i][1], 0)
            ref[i][1], _, _, _ = oracle.env_step(ref[i][0], a)
        qp = np.stack([np.array(ref[i][0].st.qpos) for i in range(n)], 1).astype(np.float32)
        qv = np.stack([np.array(ref[i][0].st.qvel) for i in range(n)], 1).astype(np.float32)
        if t % 6 == 5:       # push a slide and a finger joint past their limits so that limit rows exist as well
            qp[2, 0] = 0.5004; qp[3, 1] = -0.0006
        h.set_state_host(qp, qv)
        for i in range(n):
            s = kgo.State.from_buffer_copy(ref[i][0].st)
            for k in range(16): s.qpos[k] = float(qp[k, i])
            for k in range(15): s.qvel[k] = float(qv[k, i])
            ctrl = np.array([0, 0, 0.05, 0.3, 0.3, 0.3, 0.3, 0.3, 0.3])
            d = h.forward_debug(i, ctrl)
            if d["ncon"] > 23:
                continue                                       # (the dump holds 23 contacts)
            given = []
            for c in d["con"]:
                code = int(round(float(c[7]) * 10)); g2 = code // 1000; g1 = (code // 100) % 10
                given.append([g1, g2, float(c[0])] + [float(x) for x in c[1:7]])
            D = oracle.forward_contacts(s, ctrl, np.array(given).reshape(-1, 9))
            assert D.ncon == d["ncon"]
            mine = Handle.explicit_rows(d)
            J = np.array(D.efc_J[:D.nefc * 15]).reshape(D.nefc, 15)
            kind = {0: 1, 1: 2, 2: 3}
            assert len(mine) == D.nefc, (t, i, len(mine), D.nefc)
            used = set()
            for r in range(D.nefc):
                kd = kind[D.efc_type[r]]
                cand = [(float(np.abs(m[1] - J[r]).max()), q) for q, m in enumerate(mine) if m[0] == kd and q not in used]
                assert cand, (t, i, r, kd)
                dj, q = min(cand)
                used.add(q)
                aref, Dv = D.efc_aref[r], D.efc_D[r]
                # aref = -B J.v - K d(r) r: an fp32 dot product of 15 terms scaled by B = 2 / (dmax tc) = 105; its round-off is
                # bounded through the SUM OF THE TERMS' MAGNITUDES (cancellation), not through |aref|
                cond = 105.3 * float(np.abs(J[r] * qv[:, i].astype(np.float64)).sum())
                tol = 1e-5 * max(1.0, abs(aref)) + 16 * 2.0 ** -24 * cond
                stats[kd].append((dj / max(1.0, float(np.abs(J[r]).max())), abs(mine[q][2] - aref) / tol, abs(mine[q][3] - Dv) / Dv))
            qa = D.arr("qacc"); fc = D.arr("qfrc_constraint")
            eq = float(np.abs(qa - d["qacc"]).max() / max(1.0, np.abs(qa).max()))
            if D.ncon == 0:
                nfree += 1
                assert eq < 1e-4, (t, i, eq)
            else:
                qerr.append(eq)
    for kd, name in ((1, "tendon"), (2, "limit"), (3, "contact")):
        a = np.array(stats[kd])
        print(f"{name} rows {len(a)}: J err max {a[:, 0].max():.2e}, aref err / bound max {a[:, 1].max():.3f}, D rel err max {a[:, 2].max():.2e}")
        assert len(a) >= (400 if kd == 3 else 8), (kd, len(a))
        assert a[:, 0].max() < 2e-6 and a[:, 1].max() < 1.0 and a[:, 2].max() < 1e-5, (name, a.max(axis=0))
    qerr = np.array(qerr)
    print(f"solver on identical rows: {len(qerr)} contact states, qacc within 1e-4: {np.mean(qerr < 1e-4):.4f}, max {qerr.max():.2e}; contact-free states {nfree}")
    assert len(qerr) > 60 and nfree >= 1
    assert qerr.max() < 1e-4, (np.mean(qerr < 1e-4), qerr.max())


def test_expert_action_as_its_own_call(backend, card):
    """kg_expert_action + kg_step(action) == kg_step(action=NULL, mode): the reference's split `action = controller.get_action(obs);
    env.step(action)` (EXP:745-804) through the C ABI gives bit-identical rollouts to the fused call, in 4-d and 6-d action layouts."""
    n = 24
    shape, _, objp = _poses(card, n, 17, orient_idx=(0,), shapes=(0, 1, 15))
    _, euler, _ = _poses(card, n, 17, orient_idx=(2,), shapes=(0, 1, 15))
    for mode in (0, 1, 2):
        for dim in (4, 6):
            a, b = Handle(n, lib=backend), Handle(n, lib=backend)
            b.set_action_dim(dim)
            oa, ob = a.reset_host(shape, euler, objp), b.reset_host(shape, euler, objp)
            assert np.array_equal(oa, ob)
            for t in range(16):
                oa, ra, da = a.step_host(None, mode)
                act = b.expert_action_host(mode)
                assert act.shape == (dim, n) and (dim == 4 or not act[:2].any())
                ob, rb, db = b.step_host(act, 0)
                assert np.array_equal(act[dim - 4:], a.last_action_host()), (mode, dim, t)
                assert np.array_equal(oa, ob) and np.array_equal(ra, rb) and np.array_equal(da, db), (mode, dim, t)
            assert a.last_action_host()[0].max() > 0          # the lift phase was reached
            a.close(); b.close()


def test_expert_modes_and_rollout_success_parity(backend, oracle, card):
    """Free-running expert episodes (on-device controller) vs the oracle: the controller's actions follow the validated
    host controller on the kernel's own observations, and lift success agrees within +-2 pp (+ small-sample slack)."""
    emu = backend._name.endswith("libkg_emu.so")
    n = 48 if emu else 512
    shape, _, objp = _poses(card, n, 13, orient_idx=(0,), shapes=(0, 1))
    _, euler, _ = _poses(card, n, 13, orient_idx=(2,), shapes=(0, 1))
    for mode in ((1, 2) if emu else (0, 1, 2)):                  # naive, position-dependent, combined (EXP:610-671)
        h = Handle(n, lib=backend)
        obs = h.reset_host(shape, euler, objp)
        ref = _oracle_reset(oracle, shape, euler, objp)
        ctl = [kgo.Env() for _ in range(n)]
        for i in range(n):
            ctl[i].init_obj_x = float(obs[21, i]); ctl[i].init_dot = float(obs[81, i])
        done_k = np.zeros(n, bool); succ_k = np.zeros(n, bool); done_o = np.zeros(n, bool); succ_o = np.zeros(n, bool)
        for t in range(30):
            prev = obs.copy()
            obs, rew, done = h.step_host(None, mode)
            act = h.last_action_host()
            for i in range(n):
                if not done_k[i]:
                    a, lift = oracle.expert_action(ctl[i], prev[:, i].astype(np.float64), mode)
                    assert np.abs(a - act[:, i]).max() < 1e-5, (t, i, a, act[:, i])
                    if done[i]:
                        done_k[i] = True; succ_k[i] = rew[i] > 0
                if not done_o[i]:
                    a, _ = oracle.expert_action(ref[i][0], ref[i][1], mode)
                    ref[i][1], r, d, _ = oracle.env_step(ref[i][0], a)
                    if d:
                        done_o[i] = True; succ_o[i] = r > 0
        assert done_k.all() and done_o.all()                      # TimeLimit 30
        slack = max(0.02 * n, 3)
        assert abs(int(succ_k.sum()) - int(succ_o.sum())) <= slack, (succ_k.sum(), succ_o.sum())
        assert succ_o.sum() > 0.3 * n                             # the side grasp does lift (sanity band, expert_plots/*.npy)
        assert h.stats()["contact_overflow"] == 0


@pytest.mark.gpu
def test_launch_schedule_does_not_change_results(card):
    """The load-sorted warp-slot -> env schedule (kg_set_schedule) is a placement decision only: 12 expert RL steps with
    auto-reset give bit-identical observations, rewards, dones and states with the schedule on and off."""
    from kinovagrasping_b200 import _lib
    lib = _lib.load()
    n = 600                                           # not a multiple of the CTA's 16 slots
    shape, euler, objp = _poses(card, n, 11, orient_idx=(2,), shapes=(0, 1, 4, 8))
    runs = []
    for mode in (0, 1):
        h = Handle(n, lib=lib)
        h.set_schedule(mode)
        h.reset_host(shape, euler, objp)
        trace = []
        for _ in range(12):
            obs, rew, done = h.step_host(None, 2)
            trace.append((obs.copy(), rew.copy(), done.copy()))
        trace.append(h.get_state_host())
        runs.append(trace)
        h.close()
    for a, b in zip(*runs):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


@pytest.mark.gpu
def test_split_and_full_workspace_modes_agree(card):
    """kg_set_mode: the light-workspace kernels with their heavy lane (mode 1) and the full-workspace kernels (mode 2) run
    the same arithmetic on the same carry records; they are separate template instantiations, so fp32 rounding may differ in
    the last bits, and contact dynamics amplifies that over a free rollout.  So the two are held together in LOCK STEP: every
    RL step both start from mode 1's state (written into both, which also invalidates both collision caches) and take the
    action mode 1's on-device expert chose; a 12-step expert rollout into contact on cones and vases, where single substeps
    exceed the light workspace's 8 contacts.  Per (step, env) the 82 obs entries agree to 1e-4 for >= 98 % (median < 1e-5),
    rewards and done flags are equal for >= 99.5 %."""
    import torch
    from kinovagrasping_b200 import _lib
    lib = _lib.load()
    n = 700
    shape, euler, objp = _poses(card, n, 5, orient_idx=(2,), shapes=(0, 1, 6, 9, 10, 13))
    A = Handle(n, lib=lib); A.set_mode(1)
    B = Handle(n, lib=lib); B.set_mode(2)
    A.reset_host(shape, euler, objp); B.reset_host(shape, euler, objp)
    dev = "cuda"
    oA = torch.zeros((82, n), device=dev); oB = torch.zeros((82, n), device=dev)
    rA = torch.zeros(n, device=dev); rB = torch.zeros(n, device=dev)
    dA = torch.zeros(n, dtype=torch.int32, device=dev); dB = torch.zeros(n, dtype=torch.int32, device=dev)
    act = torch.zeros((4, n), device=dev)
    diffs = []; same = []
    for t in range(12):
        qp, qv = A.get_state_host()
        A.set_state_host(qp, qv); B.set_state_host(qp, qv)
        A.step(0, 2, oA.data_ptr(), rA.data_ptr(), dA.data_ptr(), act.data_ptr()); A.sync()
        B.step(act.data_ptr(), 0, oB.data_ptr(), rB.data_ptr(), dB.data_ptr()); B.sync()
        diffs.append((oA - oB).abs().max(dim=0).values.cpu().numpy())
        same.append(float(((rA == rB) & (dA == dB)).float().mean()))
    sa, sb = A.stats(), B.stats()
    assert sa["contact_overflow"] == 0 and sb["contact_overflow"] == 0
    assert sa["contacts"] / sa["substeps"] > 2.0                        # the rollout really is in contact
    diffs = np.stack(diffs)
    print(f"split vs full workspace, lock step: obs diff median {np.median(diffs):.2e}, within 1e-4: {(diffs <= 1e-4).mean():.4f}, "
          f"within 1e-3: {(diffs <= 1e-3).mean():.4f}, max {diffs.max():.2e}; reward/done equal: min over steps {min(same):.4f}")
    assert (diffs[0] <= 1e-4).mean() >= 0.99 and np.median(diffs[0]) < 1e-6, ((diffs[0] <= 1e-4).mean(), diffs[0].max())
    assert np.median(diffs) < 1e-5 and (diffs <= 1e-4).mean() >= 0.98, (np.median(diffs), (diffs <= 1e-4).mean())
    assert min(same) >= 0.995, same
    A.close(); B.close()


@pytest.mark.gpu
def test_heavy_lane_claims_each_substep_once_large_grid(card):
    """in split mode the heavy lane and a light warp both decide from an env's carry record, which the one
    that runs the env rewrites in the same launch.  With more CTAs than one resident wave (32768 envs, schedule off,
    contact-heavy cone / vase / 45-degree shapes in the top grasp) a heavy CTA can finish an env before its light CTA is
    scheduled; without the atomicCAS claim on subseq[env] that env was integrated twice.  Checks: (a) two identical split-mode
    runs are bit-identical (a double-run depends on CTA timing, so it shows up as non-determinism); (b) the substep counter is
    exactly envs x 15 x steps; (c) split mode agrees with full mode (no heavy lane at all) to within the last-bit differences of
    two template instantiations for the bulk of the envs, rewards included."""
    import torch
    from kinovagrasping_b200 import _lib
    lib = _lib.load()
    n, steps = 32768, 8
    shapes = (4, 5, 6, 7, 8, 9, 10, 11, 12, 13)
    shape = torch.from_numpy(np.array([shapes[i % len(shapes)] for i in range(n)], dtype=np.int32)).cuda()
    runs = []
    for mode in (1, 1, 2):
        h = Handle(n, lib=lib)
        h.set_mode(mode); h.set_schedule(0)
        obs = torch.zeros((82, n), device="cuda"); rew = torch.zeros(n, device="cuda"); done = torch.zeros(n, dtype=torch.int32, device="cuda")
        h.reset_sampled(0, shape.data_ptr(), 2, 0, 0, 77, obs.data_ptr())
        first = None
        for t in range(steps):
            h.step(0, 2, obs.data_ptr(), rew.data_ptr(), done.data_ptr())
            if t == 0:
                h.sync(); first = obs.cpu().numpy().copy()
        h.sync()
        st = h.stats()
        assert st["substeps"] == n * 15 * steps, st
        runs.append((obs.cpu().numpy().copy(), rew.cpu().numpy().copy(), done.cpu().numpy().copy()) + tuple(h.get_state_host()) + (st, first))
        h.close()
    a, b, c = runs
    for x, y in zip(a[:5], b[:5]):
        assert np.array_equal(x, y)
    assert a[5] == b[5] and np.array_equal(a[6], b[6])
    # (c): after ONE RL step the two instantiations differ by rounding only; after 8 free steps contact dynamics has amplified
    # that in some envs (test_split_and_full_workspace_modes_agree holds the two together step by step instead)
    d1 = np.abs(a[6] - c[6]).max(axis=0)
    close = np.abs(a[0] - c[0]).max(axis=0) <= 1e-3
    print(f"split vs full, 32768 envs: after 1 step within 1e-4: {(d1 <= 1e-4).mean():.4f} (median {np.median(d1):.1e}); after {steps} free steps within 1e-3: {close.mean():.4f}; rewards equal {(a[1] == c[1]).mean():.4f}")
    assert (d1 <= 1e-4).mean() >= 0.99, (d1 <= 1e-4).mean()
    assert close.mean() >= 0.85, close.mean()
    assert (a[1] == c[1]).mean() >= 0.99


@pytest.mark.gpu
def test_handles_share_the_device_model_tables(card, oracle):
    """the kernels read the model through one __constant__ symbol per device, so its tables must outlive
    any single handle.  The reference's own flow: eval_policy (MAIN:146) builds a temporary env while the training env is
    alive and drops it.  Create A, create B, step both, destroy B, keep stepping A: A's results equal those of a run that
    never saw B (bit-identical), and a third handle created afterwards still works."""
    from kinovagrasping_b200 import _lib
    lib = _lib.load()
    n = 256
    shape, euler, objp = _poses(card, n, 21, orient_idx=(2,), shapes=(0, 1, 6, 10))

    def run(with_b):
        a = Handle(n, lib=lib)
        a.reset_host(shape, euler, objp)
        out = [a.step_host(None, 2)[0].copy()]
        if with_b:
            b = Handle(64, lib=lib)
            b.reset_host(shape[:64], euler[:, :64], objp[:, :64])
            b.step_host(None, 0)
            out.append(a.step_host(None, 2)[0].copy())
            b.close()
        else:
            out.append(a.step_host(None, 2)[0].copy())
        for _ in range(3):
            out.append(a.step_host(None, 2)[0].copy())
        a.close()
        return out
    ref, got = run(False), run(True)
    for x, y in zip(ref, got):
        assert np.array_equal(x, y)
    c = Handle(8, lib=lib)                              # tables were freed with the last handle and are uploaded again
    o = c.reset_host(shape[:8], euler[:, :8], objp[:, :8])
    r = _oracle_reset(oracle, shape[:8], euler[:, :8], objp[:, :8])
    assert max(_obs_err(o[:, i].astype(np.float64), r[i][1]).max() for i in range(8)) < 1e-5
    c.close()


@pytest.mark.gpu
def test_cuda_graph_replay_does_not_change_results(card):
    """kg_set_graph: a step replayed from the captured CUDA graph of its kernel chain is the same arithmetic as the
    individually launched chain -- bit-identical obs / reward / done / state over expert and caller-driven steps, with the
    argument sets changing between calls (expert vs action tensor, outputs on / off, a setter in between: each needs its own
    graph instance, keyed by kg_api.cu's KgGraphKey) and auto-resets in between."""
    import torch
    from kinovagrasping_b200 import _lib
    lib = _lib.load()
    n = 2048
    shape = torch.from_numpy((np.arange(n) % 2).astype(np.int32)).cuda()
    act = torch.zeros((4, n), device="cuda"); act[1:] = 0.45
    runs = []
    for graph in (0, 1):
        h = Handle(n, lib=lib)
        h.set_graph(graph)
        h.set_stream(torch.cuda.current_stream().cuda_stream)
        obs = torch.zeros((82, n), device="cuda"); rew = torch.zeros(n, device="cuda"); done = torch.zeros(n, dtype=torch.int32, device="cuda")
        h.reset_sampled(0, shape.data_ptr(), 0, 1, 0, 5, obs.data_ptr())
        trace = []
        for k in range(10):
            if k == 6:
                h.set_max_episode_steps(8)
            if k % 3 == 2:
                h.step(act.data_ptr(), 0, obs.data_ptr(), rew.data_ptr(), done.data_ptr())
            elif k % 3 == 1:
                h.step(0, 2, 0, rew.data_ptr(), done.data_ptr())
            else:
                h.step(0, 2, obs.data_ptr(), rew.data_ptr(), done.data_ptr())
            h.sync()
            trace += [obs.cpu().numpy().copy(), rew.cpu().numpy().copy(), done.cpu().numpy().copy()]
            h.reset_sampled(done.data_ptr(), 0, 0, 1, 0, 5, 0)
        trace += list(h.get_state_host())
        runs.append(trace)
        h.close()
    for x, y in zip(*runs):
        assert np.array_equal(x, y)


@pytest.mark.gpu
def test_dependent_launch_does_not_change_results(card):
    """kg_set_dependent_launch: the chain kernels launched with programmatic stream serialisation (their launch overlaps the
    predecessor, griddepcontrol.wait orders the data) give bit-identical obs / reward / done / state -- graph replay and
    individual launches, split and full workspace mode, 15 substeps and frame_skip 1, contact-heavy shapes included."""
    import torch
    from kinovagrasping_b200 import _lib
    lib = _lib.load()
    n = 4096
    shape = torch.from_numpy((np.arange(n) % 14).astype(np.int32)).cuda()
    for graph, mode, fs in ((1, 0, 15), (0, 0, 15), (1, 2, 15), (1, 0, 1)):
        runs = []
        for pdl in (0, 1):
            h = Handle(n, lib=lib)
            h.set_graph(graph); h.set_mode(mode); h.set_frame_skip(fs); h.set_dependent_launch(pdl)
            h.set_stream(torch.cuda.current_stream().cuda_stream)
            obs = torch.zeros((82, n), device="cuda"); rew = torch.zeros(n, device="cuda"); done = torch.zeros(n, dtype=torch.int32, device="cuda")
            h.reset_sampled(0, shape.data_ptr(), 3, 1, 0, 11, obs.data_ptr())
            trace = []
            for k in range(14 if fs == 15 else 60):
                h.step(0, 2, obs.data_ptr(), rew.data_ptr(), done.data_ptr())
                h.sync()
                trace += [obs.cpu().numpy().copy(), rew.cpu().numpy().copy(), done.cpu().numpy().copy()]
                h.reset_sampled(done.data_ptr(), 0, 3, 1, 0, 11, 0)
            trace += list(h.get_state_host())
            st = h.stats()
            assert st["contact_overflow"] == 0 and st["mpr_pairs"] > 0
            runs.append(trace)
            h.close()
        for x, y in zip(*runs):
            assert np.array_equal(x, y), (graph, mode, fs)


@pytest.mark.gpu
def test_pipeline_groups_do_not_change_results(card):
    """kg_set_pipeline cuts the batch into env groups whose kernel chains run on separate internal streams; an env's
    arithmetic does not depend on its group, so 1, 2 and 3 groups give bit-identical observations, rewards, dones and
    states -- checked at the benchmark's batch size (4096 envs, CubeS/CubeB, 8 expert RL steps with auto-reset of the
    finished envs) together with determinism of a repeated run."""
    import torch
    from kinovagrasping_b200 import _lib
    lib = _lib.load()
    n = 4096
    shape = torch.from_numpy((np.arange(n) % 2).astype(np.int32)).cuda()
    runs = []
    for groups in (1, 2, 3, 2):
        h = Handle(n, lib=lib)
        h.set_pipeline(groups)
        obs = torch.zeros((82, n), device="cuda"); rew = torch.zeros(n, device="cuda"); done = torch.zeros(n, dtype=torch.int32, device="cuda")
        h.reset_sampled(0, shape.data_ptr(), 0, 1, 0, 99, obs.data_ptr())
        trace = [obs.cpu().numpy().copy()]
        for _ in range(8):
            h.step(0, 2, obs.data_ptr(), rew.data_ptr(), done.data_ptr())
            h.sync()
            trace += [obs.cpu().numpy().copy(), rew.cpu().numpy().copy(), done.cpu().numpy().copy()]
            h.reset_sampled(done.data_ptr(), 0, 0, 1, 0, 99, 0)
        trace += list(h.get_state_host())
        st = h.stats()
        assert st["contact_overflow"] == 0
        # the collision caches do their job: before the grasp closes nearly every near pair is settled by its cached axis
        # (without them all ~13 near pairs of an env go to the narrowphase in every substep)
        assert st["mpr_pairs"] / st["substeps"] < 2.5, st
        runs.append(trace)
        h.close()
    for other in runs[1:]:
        for x, y in zip(runs[0], other):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("mode", [1, 2])
def test_no_contact_is_dropped_above_the_full_workspace(backend, oracle, card, mode):
    """MuJoCo never drops a contact for this model, the round-1 kernels did above 23.  A hand pressed flat on the
    floor with folded fingers has 25+ contacts (plane lanes of the palm and all six finger geoms + the object): more than
    the light (8) AND the full (23) workspace hold.  One physics substep from that state, in split mode and in full mode:
    the heavy lane of the launch runs it with rows for all 54 possible contacts -- kg_get_stats reports no overflow, the
    contact count equals the oracle's, and the next velocities agree with the oracle's step."""
    emu = backend._name.endswith("libkg_emu.so")
    if emu and mode == 2:
        pytest.skip("the emulation has one workspace only")
    n = 4
    shape = np.array([9, 9, 0, 0], dtype=np.int32)
    euler = np.tile(np.array([[-1.57], [0.0], [-1.57]], dtype=np.float32), (1, n))
    objp = np.tile(np.array([[0.0], [0.0], [0.03]], dtype=np.float32), (1, n))
    objp[:, 3] = card.pose("train")[0, 0, 7, 0:3]      # env 3: an ordinary start (object standing on the floor)
    h = Handle(n, lib=backend)
    h.set_mode(mode); h.set_frame_skip(1); h.set_max_episode_steps(10 ** 6)
    h.reset_host(shape, euler, objp)
    qp, qv = h.get_state_host()
    qp[0], qp[1], qp[2] = 0.0, -0.04, -0.08
    qp[3] = qp[5] = qp[7] = 1.2; qp[4] = qp[6] = qp[8] = 0.6
    qp[0:9, 3] = 0.0                                   # env 3: open hand above the object (an ordinary substep next to the heavy ones)
    h.set_state_host(qp, qv)
    h.stats()
    obs, rew, done = h.step_host(np.zeros((4, n), dtype=np.float32), 0)
    st = h.stats()
    assert st["contact_overflow"] == 0 and st["substeps"] == n
    qp2, qv2 = h.get_state_host()
    ncons = []
    for i in range(n):
        s = oracle.state(int(shape[i]), euler[:, i].astype(np.float64), objp[:, i].astype(np.float64))
        for k in range(16): s.qpos[k] = float(qp[k, i])
        for k in range(15): s.qvel[k] = float(qv[k, i])
        import ctypes as C
        D0 = oracle.forward(s)
        c = np.zeros(9); a = np.zeros(4); dp = C.POINTER(C.c_double)
        kgo.lib().kgo_ctrl_from_action(C.byref(D0), a.ctypes.data_as(dp), c.ctypes.data_as(dp))
        D = oracle.step(s, c)
        ncons.append(D.ncon)
        ref_v = np.array(s.qvel)
        assert np.abs(qv2[:, i] - ref_v).max() < 2e-3 * max(1.0, np.abs(ref_v).max()), (i, D.ncon, np.abs(qv2[:, i] - ref_v).max(), np.abs(ref_v).max())
    assert max(ncons[:3]) > 23 and ncons[3] <= 8, ncons
    assert st["contacts"] == sum(ncons), (st, ncons)
    h.close()


def test_global_frame_obs_parity(backend, oracle, card):
    """_get_obs(state_rep="global"): the 74-value world-frame layout written next to the local one (kg_set_obs_global)
    against the oracle (itself pinned to the reference's own _get_obs in test_oracle_golden), after reset for all shapes and
    orientations and after two expert RL steps; the entries the two layouts share are the same numbers."""
    n = 42
    shape, euler, objp = _poses(card, n, 4)
    h = Handle(n, lib=backend)
    h.enable_obs_global()
    obs = h.reset_host(shape, euler, objp)
    g = h.obs_global_host()
    ref = _oracle_reset(oracle, shape, euler, objp)
    for i in range(n):
        og = oracle.get_obs(ref[i][0].st, ref[i][0].d, global_rep=True)
        assert _obs_err(g[:, i].astype(np.float64), og).max() < 1e-5
    assert np.array_equal(g[24:67], obs[24:67]) and np.array_equal(g[67:74], obs[75:82])
    for _ in range(2):
        obs, rew, done = h.step_host(None, 2)
    g = h.obs_global_host()
    assert np.array_equal(g[24:67], obs[24:67]) and np.array_equal(g[67:74], obs[75:82])
    assert np.abs(g[21:24] - h.get_state_host()[0][9:12]).max() < 0.06      # the object geom sits near the free joint's position
    h.close()
