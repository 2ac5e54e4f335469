"""GPU-resident episode replay with n-step window sampling: the reference's `ReplayBuffer_Queue`
(utils.py:9-460) restated as one preallocated ring of episodes in device memory.

Reference semantics kept (each cited where it is implemented):
  * episode-structured FIFO: `add` appends to the episode in progress, a `done` transition closes it and the oldest
    episode is evicted once `replay_ep_num >= max_episode` (utils.py:34-79); `replace` rewrites the last stored
    reward/done of the episode in progress and closes it, evicting at `>` (utils.py:309-343);
  * `add_episode` start/finish bookkeeping of `episodes` / `timesteps_count` (utils.py:81-101);
  * `sample_batch_nstep(batch_size)`: `batch_size` EPISODES drawn uniformly from [0, replay_ep_num-1), and from each
    `ceiling = len - n` windows of n consecutive steps (`ceiling-1` random starts in [0, ceiling) plus the last window),
    so the result has sum(ceiling) rows (utils.py:240-306).  The draws use `np.random` in the reference's own call
    order, so a seeded reference run and a seeded run of this class return identical batches (tests/test_learner.py).
What changes: storage is `[episode][t][dim]` fp32 tensors on the device (no Python lists, no per-sample H2D), windows
are gathered by one index_select per field, and the vectorised environment feeds it with `add_batch` (one scatter per
step for all N environments).  `sample_windows` is the fixed-shape, all-on-device sampler used under CUDA graphs."""
import numpy as np
import torch


class ReplayBuffer_Queue:
    def __init__(self, state_dim, action_dim, max_episode=10000, n_steps=5, max_episode_len=32, device=None, n_envs=0):
        self.state_dim, self.action_dim = int(state_dim), int(action_dim)
        self.max_episode = int(max_episode)
        self.n_steps = int(n_steps)
        self.device = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.T = int(max_episode_len)
        self.cap = self.max_episode + 2            # complete episodes + the one in progress + one spare row
        self.size = 0                              # transitions ever added (utils.py:12)
        self.episodes_count = 0                    # episodes ever closed
        self.replay_ep_num = 0                     # complete episodes currently held
        self.episodes = []
        self.timesteps_count = 0
        self.orientation_indexes = []
        self.head = 0                              # physical row of logical episode 0
        self.lengths = np.zeros(self.cap, dtype=np.int64)   # host mirror of the per-row lengths
        self._alloc(self.T)
        # staging for the vectorised feeder (episodes in progress, one per environment)
        self.n_envs = int(n_envs)
        if self.n_envs:
            f = dict(dtype=torch.float32, device=self.device)
            self.st_state = torch.zeros((self.n_envs, self.T, self.state_dim), **f)
            self.st_action = torch.zeros((self.n_envs, self.T, self.action_dim), **f)
            self.st_next = torch.zeros((self.n_envs, self.T, self.state_dim), **f)
            self.st_reward = torch.zeros((self.n_envs, self.T), **f)
            self.st_notdone = torch.zeros((self.n_envs, self.T), **f)
            self.st_len = torch.zeros(self.n_envs, dtype=torch.long, device=self.device)
            self.st_closed = torch.zeros(self.n_envs, dtype=torch.bool, device=self.device)
            self._env_ids = torch.arange(self.n_envs, device=self.device)
        self.len_dev = torch.zeros(self.cap, dtype=torch.long, device=self.device)
        self._host_len_dirty = False
        # device copy of (head, complete episodes - 1) for the device sampler: a captured CUDA graph replays sample_windows
        # without running this Python, so the ring's bookkeeping has to be data, not constants (publish_meta)
        self._meta_dev = torch.zeros(2, dtype=torch.long, device=self.device)
        self._meta_host = (-1, -1)

    def publish_meta(self):
        """Refresh the device copy of the ring bookkeeping; call before replaying a graph that contains sample_windows."""
        cur = (int(self.head), int(self.replay_ep_num) - 1)
        if cur != self._meta_host:
            self._meta_dev.copy_(torch.tensor(cur, dtype=torch.long))
            self._meta_host = cur
        return self._meta_dev

    # ------------------------------------------------------------------ storage
    def _alloc(self, T):
        """(Re)allocate the ring for episodes of up to T steps.  Growing replaces every storage AND staging tensor, so
        anything that captured their addresses (a CUDA graph around sample_windows) is stale afterwards:
        `storage_version` is bumped and GraphedUpdates re-captures when it changes."""
        f = dict(dtype=torch.float32, device=self.device)

        def grown(names, lead):
            new = [torch.zeros((lead, T, self.state_dim), **f), torch.zeros((lead, T, self.action_dim), **f),
                   torch.zeros((lead, T, self.state_dim), **f), torch.zeros((lead, T), **f), torch.zeros((lead, T), **f)]
            if hasattr(self, names[0]):
                for nm, nw in zip(names, new):
                    old = getattr(self, nm)
                    nw[:, :old.shape[1]] = old
            for nm, nw in zip(names, new):
                setattr(self, nm, nw)
        grown(("state", "action", "next_state", "reward", "not_done"), self.cap)
        if getattr(self, "n_envs", 0) and hasattr(self, "st_state"):
            grown(("st_state", "st_action", "st_next", "st_reward", "st_notdone"), self.n_envs)
        self.T = T
        self.storage_version = getattr(self, "storage_version", -1) + 1

    def _sync_lengths(self):
        if self._host_len_dirty:                   # the vectorised feeder only updates the device copy
            self.lengths[:] = self.len_dev.cpu().numpy()
            self._host_len_dirty = False

    def _row(self, logical):
        return (self.head + logical) % self.cap

    def _t(self, x, dim):
        return torch.as_tensor(np.asarray(x, dtype=np.float32).reshape(dim), dtype=torch.float32).to(self.device, non_blocking=True)

    # ------------------------------------------------------------------ reference single-transition API
    def add(self, state, action, next_state, reward, done):
        """utils.py:34-60.  (The reference's lists are unbounded: after remove_episode(-1) the next episode is appended to
        the short one, so a row can outgrow max_episode_len -- the ring then grows, see _alloc.)"""
        self._sync_lengths()
        row = self._row(self.replay_ep_num)
        t = int(self.lengths[row])
        if t >= self.T:
            self._alloc(2 * self.T)
        self.state[row, t] = self._t(state, self.state_dim)
        self.action[row, t] = self._t(action, self.action_dim)
        self.next_state[row, t] = self._t(next_state, self.state_dim)
        self.reward[row, t] = float(reward)
        self.not_done[row, t] = 1.0 - float(done)
        self.lengths[row] = t + 1
        self.len_dev[row] = t + 1
        self.size += 1
        self.timesteps_count += 1
        if done:
            self.episodes_count += 1
            self.replay_ep_num += 1
            self._open_row()
            if self.replay_ep_num >= self.max_episode:
                self.remove_episode()

    def _open_row(self):
        row = self._row(self.replay_ep_num)
        self.lengths[row] = 0
        self.len_dev[row] = 0

    def remove_episode(self, idx=0):
        """utils.py:62-79: `list.pop(idx)` on the per-episode lists, whose entries are the replay_ep_num closed episodes
        followed by the one in progress.  idx 0 is the FIFO eviction.  main_DDPGfD.py:470-471 calls it with -1 for an
        episode too short to train on -- which pops the (empty) in-progress entry, so the short episode itself becomes the
        episode in progress again and the next episode's transitions are appended to it.  Reproduced as is."""
        self._sync_lengths()
        n_rows = self.replay_ep_num + 1
        i = idx + n_rows if idx < 0 else idx
        if not 0 <= i < n_rows:
            raise IndexError("pop index out of range")
        if i == 0:
            self.lengths[self._row(0)] = 0
            self.len_dev[self._row(0)] = 0
            self.head = (self.head + 1) % self.cap
        else:
            for j in range(i, n_rows - 1):          # close the gap (rare: only the ends are ever popped by the reference's callers)
                a, b = self._row(j), self._row(j + 1)
                for x in (self.state, self.action, self.next_state, self.reward, self.not_done):
                    x[a] = x[b]
                self.lengths[a] = self.lengths[b]
            last = self._row(n_rows - 1)
            self.lengths[last] = 0
            self.len_dev.copy_(torch.from_numpy(self.lengths))
        if self.episodes:                            # (only maintained by callers that use add_episode)
            self.episodes.pop(idx)
        if len(self.orientation_indexes) > 0:
            self.orientation_indexes.pop(idx)
        self.replay_ep_num -= 1

    def add_episode(self, start):
        """utils.py:81-101."""
        if start:
            if self.replay_ep_num == 0:
                self.episodes.append([self.timesteps_count])
            else:
                self.episodes[-1].append(self.timesteps_count)
        else:
            if self.replay_ep_num == 0:
                self.episodes.append(self.timesteps_count)
                self.episodes.append([])
                self.timesteps_count = 0
            else:
                self.episodes[-1].append(self.timesteps_count)
                self.episodes.append([])
                self.timesteps_count = 0

    def add_orientation_idx_to_replay(self, idx):
        self.orientation_indexes.append(idx)

    def replace(self, reward, done):
        """utils.py:309-343: lifted-and-done rewrites the last stored step of the episode in progress."""
        if not done:
            print("Can only replace last time step of the latest episode")
            raise ValueError
        self._sync_lengths()
        row = self._row(self.replay_ep_num)
        idx = int(self.lengths[row]) - 1
        old_reward = float(self.reward[row, idx])
        self.reward[row, idx] = float(reward)
        self.not_done[row, idx] = 1.0 - float(done)
        self.episodes_count += 1
        self.replay_ep_num += 1
        self._open_row()
        if self.replay_ep_num > self.max_episode:
            self.remove_episode()
        return old_reward

    def episode_length(self, logical):
        self._sync_lengths()
        return int(self.lengths[self._row(logical)])

    def episode(self, logical):
        """(state, action, next_state, reward, not_done) views of one stored episode."""
        row, n = self._row(logical), self.episode_length(logical)
        return self.state[row, :n], self.action[row, :n], self.next_state[row, :n], self.reward[row, :n], self.not_done[row, :n]

    # ------------------------------------------------------------------ sampling
    def _gather(self, rows, starts):
        rows = torch.as_tensor(np.asarray(rows, dtype=np.int64), device=self.device)
        starts = torch.as_tensor(np.asarray(starts, dtype=np.int64), device=self.device)
        return self._gather_dev(rows, starts)

    def _gather_dev(self, rows, starts):
        t = starts[:, None] + torch.arange(self.n_steps, device=self.device)[None, :]
        flat = (rows[:, None] * self.T + t).reshape(-1)
        W = rows.shape[0]

        def g(x):
            return x.reshape(self.cap * self.T, *x.shape[2:]).index_select(0, flat).reshape(W, self.n_steps, *x.shape[2:])
        return g(self.state), g(self.action), g(self.next_state), g(self.reward), g(self.not_done)

    def sample_batch_nstep(self, batch_size, num_ts_from_ep=5):
        """utils.py:240-306, same np.random call order: one randint(replay_ep_num-1, size=batch_size), then per
        episode (ceiling-1) x randint(ceiling)."""
        if batch_size < 1:
            e = torch.zeros(0, device=self.device)
            return e, e, e, e, e
        episode_idx_arr = np.random.randint(self.replay_ep_num - 1, size=batch_size)
        rows, starts = [], []
        for idx in episode_idx_arr:
            episode_len = self.episode_length(int(idx))
            ceiling = episode_len - self.n_steps
            if ceiling < 1:
                raise ValueError(f"episode {idx} has {episode_len} steps: too short for {self.n_steps}-step windows "
                                 "(callers drop such episodes, main_DDPGfD.py:470-471)")
            for _ in range(ceiling - 1):
                starts.append(np.random.randint(ceiling)); rows.append(self._row(int(idx)))
            starts.append(ceiling); rows.append(self._row(int(idx)))
        return self._gather(rows, starts)

    def sample_windows(self, n_windows, generator=None):
        """Fixed-shape device sampler with the same per-window distribution as sample_batch_nstep: an episode is drawn with
        probability proportional to its window count `ceiling` (the reference emits `ceiling` rows per drawn episode),
        then the window is the last one with probability 1/ceiling and uniform over [0, ceiling) otherwise.
        No host synchronisation: usable inside a captured CUDA graph."""
        capturing = self.device.type == "cuda" and torch.cuda.is_current_stream_capturing()
        meta = self._meta_dev if capturing else self.publish_meta()
        logical = torch.arange(self.cap, device=self.device)
        rows_all = (meta[0] + logical) % self.cap
        ceil_all = (self.len_dev[rows_all] - self.n_steps).clamp(min=0).float()
        ceil_all = torch.where(logical < meta[1], ceil_all, torch.zeros_like(ceil_all))
        pick = torch.multinomial(ceil_all, n_windows, replacement=True, generator=generator)
        rows = rows_all[pick]
        ceiling = ceil_all[pick]
        u = torch.rand(n_windows, 2, device=self.device, generator=generator)
        last = u[:, 0] * ceiling < 1.0
        starts = torch.where(last, ceiling, torch.floor(u[:, 1] * ceiling)).long()
        return self._gather_dev(rows, starts)

    # ------------------------------------------------------------------ vectorised feeder
    def add_batch(self, state, action, next_state, reward, done, record=None, replace=None, defer=False):
        """One step of all N environments (tensors on the device, env-major [N, dim]).
        record[i]  : store this transition (default all).  The rollout loops do not store the lifting steps
                     (expert_data.py:794, main_DDPGfD.py:451-456).
        replace[i] : lifted-and-done: rewrite the last stored reward with `reward[i]` and close the episode
                     (utils.py:309-343).
        A stored `done` transition closes the episode as in `add`.  Closed episodes with len - n_steps <= 1 are dropped
        (main_DDPGfD.py:470-471); the others are committed to the ring in env order, evicting the oldest when full.
        defer=False: commit right away and return the number of episodes committed (one host synchronisation).
        defer=True : only stage -- no host synchronisation at all, every update is a masked device write; closed episodes
                     wait in their env's staging row (which ignores further transitions) until `flush()` commits them.  This
                     is what the lock-step rollout uses: an env closes at most one episode per sweep."""
        N = self.n_envs
        dev = self.device
        done = done.to(dev).bool()
        record = torch.ones(N, dtype=torch.bool, device=dev) if record is None else record.to(dev).bool()
        replace = torch.zeros(N, dtype=torch.bool, device=dev) if replace is None else replace.to(dev).bool()
        open_ = ~self.st_closed
        record = record & open_
        e = self._env_ids
        t = self.st_len.clamp(max=self.T - 1)
        m1 = record[:, None]
        self.st_state[e, t] = torch.where(m1, state, self.st_state[e, t])
        self.st_action[e, t] = torch.where(m1, action, self.st_action[e, t])
        self.st_next[e, t] = torch.where(m1, next_state, self.st_next[e, t])
        self.st_reward[e, t] = torch.where(record, reward.float(), self.st_reward[e, t])
        self.st_notdone[e, t] = torch.where(record, 1.0 - done.float(), self.st_notdone[e, t])
        self.st_len += record.long()
        rep = replace & open_ & (self.st_len > 0)
        last = (self.st_len - 1).clamp(min=0)
        self.st_reward[e, last] = torch.where(rep, reward.float(), self.st_reward[e, last])
        self.st_notdone[e, last] = torch.where(rep, torch.zeros_like(self.st_reward[e, last]), self.st_notdone[e, last])
        self.st_closed |= (record & done) | rep
        return 0 if defer else self.flush()

    def flush(self):
        """Commit every closed staged episode to the ring (env order), dropping those too short to train on; returns the
        number committed.  The one place of the vectorised feeder that talks to the host (the ring's bookkeeping)."""
        dev = self.device
        closing = self.st_closed
        ids = self._env_ids[closing]
        if ids.numel() == 0:
            return 0
        lens = self.st_len[ids]
        self.size += int(lens.sum())
        keep = (lens - self.n_steps) > 1
        ids, lens = ids[keep], lens[keep]
        k = int(ids.numel())
        if k:
            if k > self.max_episode:
                ids, lens, k = ids[-self.max_episode:], lens[-self.max_episode:], self.max_episode
            evict = max(0, self.replay_ep_num + k - (self.max_episode - 1))
            if evict:
                self.head = (self.head + evict) % self.cap
                self.replay_ep_num -= evict
            rows = (self.head + self.replay_ep_num + torch.arange(k, device=dev)) % self.cap
            self.state[rows] = self.st_state[ids]
            self.action[rows] = self.st_action[ids]
            self.next_state[rows] = self.st_next[ids]
            self.reward[rows] = self.st_reward[ids]
            self.not_done[rows] = self.st_notdone[ids]
            self.len_dev[rows] = lens
            self._host_len_dirty = True
            self.replay_ep_num += k
            self.episodes_count += k
            self._open_row()
        self.st_len.masked_fill_(closing, 0)
        self.st_closed.zero_()
        return k

    def abandon(self, mask):
        """Forget the staged, still OPEN episodes of the masked environments (time-limit without a stored done)."""
        self.st_len.masked_fill_(mask.to(self.device).bool() & ~self.st_closed, 0)

    # ------------------------------------------------------------------ persistence (utils.py:345-460 layout)
    def save_replay_buffer(self, save_filepath):
        """utils.py:345-367: eight np.save files with the reference's names; the per-episode fields are object arrays of
        per-episode lists, closed episodes followed by the entry of the episode in progress.  Returns the directory WITH
        a trailing '/', as the reference does (its loader concatenates file names onto it)."""
        import os
        os.makedirs(save_filepath, exist_ok=True)
        save_filepath += "/"
        n = self.replay_ep_num

        def obj(field):
            a = np.empty(n + 1, dtype=object)
            for i in range(n + 1):
                row, ln = self._row(i), self.episode_length(i)
                a[i] = [np.asarray(v) for v in field[row, :ln].cpu().numpy()]
            return a

        def ragged(lst):
            a = np.empty(len(lst), dtype=object)
            for i, v in enumerate(lst):
                a[i] = v
            return a
        np.save(save_filepath + "state", obj(self.state), allow_pickle=True)
        np.save(save_filepath + "action", obj(self.action), allow_pickle=True)
        np.save(save_filepath + "next_state", obj(self.next_state), allow_pickle=True)
        np.save(save_filepath + "reward", obj(self.reward), allow_pickle=True)
        np.save(save_filepath + "not_done", obj(self.not_done), allow_pickle=True)
        np.save(save_filepath + "episodes", ragged(self.episodes), allow_pickle=True)
        np.save(save_filepath + "episodes_info", [self.max_episode, self.size, self.episodes_count, self.replay_ep_num])
        np.save(save_filepath + "orientation_indexes", ragged(self.orientation_indexes), allow_pickle=True)
        return save_filepath

    def _set_capacity(self, max_episode):
        """Re-seat the ring for a larger max_episode (store_saved_data_into_replay adds the file's to ours)."""
        self._sync_lengths()
        k = self.replay_ep_num + 1
        order = [self._row(i) for i in range(k)]
        old = (self.state, self.action, self.next_state, self.reward, self.not_done)
        old_len = self.lengths[order].copy()
        self.max_episode = int(max_episode)
        self.cap = self.max_episode + 2
        for nm in ("state", "action", "next_state", "reward", "not_done"):
            delattr(self, nm)
        self.lengths = np.zeros(self.cap, dtype=np.int64)
        self.len_dev = torch.zeros(self.cap, dtype=torch.long, device=self.device)
        self.head = 0
        self._alloc(self.T)
        idx = torch.as_tensor(order, dtype=torch.long, device=self.device)
        for o, nm in zip(old, ("state", "action", "next_state", "reward", "not_done")):
            getattr(self, nm)[:k] = o[idx]
        self.lengths[:k] = old_len
        self.len_dev.copy_(torch.from_numpy(self.lengths))
        self._meta_host = (-1, -1)

    def store_saved_data_into_replay(self, filepath):
        """utils.py:369-460 for a buffer that is still empty (the only way main_DDPGfD.py uses it, MAIN:1186): the saved
        per-episode lists become this buffer's rows, max_episode / size / episodes_count / replay_ep_num are ADDED to, and
        the trailing empty entry is popped with remove_episode(-1) -- which also takes 1 off replay_ep_num, so the last
        saved episode ends up as the episode in progress (reference behaviour, kept).  Loading into a non-empty buffer
        (the reference's list.extend path leaves an empty row in the middle of its lists and then samples it) appends the
        saved episodes one by one instead."""
        import os
        if filepath is None or os.path.isdir(filepath) is False:
            print("Replay buffer not found!! filepath: ", filepath)
            raise SystemExit                       # the reference calls quit()
        ld = lambda name: np.load(filepath + name + ".npy", allow_pickle=True)
        st, ac, ns, rw, nd = (ld(k).astype("object") for k in ("state", "action", "next_state", "reward", "not_done"))
        episodes = ld("episodes").astype("object")
        info = ld("episodes_info")
        if os.path.isfile(filepath + "orientation_indexes.npy"):
            self.orientation_indexes = ld("orientation_indexes").astype("object").tolist()
        if self.size != 0:
            for e in range(len(st)):
                if len(st[e]) == 0:
                    continue
                self.add_episode(1)
                for t in range(len(st[e])):
                    self.add(st[e][t], ac[e][t], ns[e][t], rw[e][t], 1.0 - float(nd[e][t]) if t == len(st[e]) - 1 else 0.0)
                if float(nd[e][-1]) != 0.0:          # an episode saved without a stored done: close it explicitly
                    self.replace(float(rw[e][-1]), True)
                self.add_episode(0)
            return ""
        rows = len(st)
        longest = max([len(e) for e in st] + [1])
        self._set_capacity(max(self.max_episode + int(info[0]), rows))
        if longest > self.T:
            self._alloc(longest)
        for e in range(rows):
            ln = len(st[e])
            if ln:
                self.state[e, :ln] = torch.as_tensor(np.asarray(list(st[e]), dtype=np.float32).reshape(ln, self.state_dim)).to(self.device)
                self.action[e, :ln] = torch.as_tensor(np.asarray(list(ac[e]), dtype=np.float32).reshape(ln, self.action_dim)).to(self.device)
                self.next_state[e, :ln] = torch.as_tensor(np.asarray(list(ns[e]), dtype=np.float32).reshape(ln, self.state_dim)).to(self.device)
                self.reward[e, :ln] = torch.as_tensor(np.asarray(list(rw[e]), dtype=np.float32).reshape(ln)).to(self.device)
                self.not_done[e, :ln] = torch.as_tensor(np.asarray(list(nd[e]), dtype=np.float32).reshape(ln)).to(self.device)
            self.lengths[e] = ln
        self.len_dev.copy_(torch.from_numpy(self.lengths))
        self.episodes = episodes.tolist()
        self.size += int(info[1])
        self.episodes_count += int(info[2])
        self.replay_ep_num += int(info[3])
        if rows != self.replay_ep_num + 1:
            raise ValueError(f"{filepath}: {rows} per-episode entries for replay_ep_num {self.replay_ep_num} (expected closed episodes + 1)")
        if int(self.lengths[self._row(rows - 1)]) == 0:      # "if there is a trailing empty entry [], remove it"
            self.remove_episode(-1)
        return ""
