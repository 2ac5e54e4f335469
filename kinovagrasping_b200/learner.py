"""DDPGfD learner on the GPU: the reference's `DDPGfD` (DDPGfD.py:15-382) with the same constructor, `select_action`,
`train_batch`, `save` / `load` (the four `torch.save` files `<name>_{actor,critic,actor_optimizer,critic_optimizer}`), so
main_DDPGfD.py's call sites stay unchanged.

Same mathematics (checked against the reference's own train_batch on identical batches, tests/test_learner.py):
  actor 82-400-300-4, ReLU, max_action*sigmoid (DDPG:15-32); critic 86-400-300-1 (DDPG:35-50); Kaiming-uniform weights;
  critic loss = MSE(Q(s0,a0), r0 + g*Q'(s1, pi'(s1))) + 0.5 * MSE(Q(s0,a0), sum_i g^i r_i + g^n Q'(s_n, pi'(s_n)))
  (DDPG:273-340; note that `not_done` is sampled but never used by the reference); actor loss = -mean Q(s, pi(s)) over ALL n
  window states (DDPG:350); Adam lr 1e-4 (actor) / 1e-3 with weight_decay 1e-4 (critic) (DDPG:57,61); Polyak tau on every
  `network_repl_freq`-th iteration only (DDPG:360-366).
What changes: every network's parameters are views into ONE flat fp32 buffer, so the data-parallel gradient exchange is a
single NCCL all-reduce per backward over NVLink (155 401 + 154 704 values); `train_on_batch` is free of host
synchronisation (losses stay on the device) so an update can be captured in a CUDA graph; batched `select_actions`
serves the vectorised env without the per-step H2D/D2H of DDPG:71-73."""
import copy

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class Actor(nn.Module):
    def __init__(self, state_dim, action_dim, max_action):
        super().__init__()
        self.l1 = nn.Linear(state_dim, 400)
        self.l2 = nn.Linear(400, 300)
        self.l3 = nn.Linear(300, action_dim)
        for l in (self.l1, self.l2, self.l3):
            torch.nn.init.kaiming_uniform_(l.weight, a=0, mode="fan_in", nonlinearity="leaky_relu")
        self.max_action = max_action

    def forward(self, state):
        a = F.relu(self.l1(state))
        a = F.relu(self.l2(a))
        return self.max_action * torch.sigmoid(self.l3(a))


class Critic(nn.Module):
    def __init__(self, state_dim, action_dim):
        super().__init__()
        self.l1 = nn.Linear(state_dim + action_dim, 400)
        self.l2 = nn.Linear(400, 300)
        self.l3 = nn.Linear(300, 1)
        for l in (self.l1, self.l2, self.l3):
            torch.nn.init.kaiming_uniform_(l.weight, a=0, mode="fan_in", nonlinearity="leaky_relu")

    def forward(self, state, action):
        q = F.relu(self.l1(torch.cat([state, action], -1)))
        q = F.relu(self.l2(q))
        return self.l3(q)


def flatten_parameters(modules):
    """Re-seat every parameter of the given modules as a view into ONE flat buffer (and its .grad into one flat gradient
    buffer); returns (flat_param, flat_grad, [(offset, numel) per module])."""
    if isinstance(modules, nn.Module):
        modules = [modules]
    ps = [p for m in modules for p in m.parameters()]
    n = sum(p.numel() for p in ps)
    flat = torch.zeros(n, dtype=ps[0].dtype, device=ps[0].device)
    grad = torch.zeros_like(flat)
    o = 0
    spans = []
    for m in modules:
        start = o
        for p in m.parameters():
            k = p.numel()
            flat[o:o + k].copy_(p.data.reshape(-1))
            p.data = flat[o:o + k].view_as(p.data)
            p.grad = grad[o:o + k].view_as(p.data)
            o += k
        spans.append((start, o - start))
    return flat, grad, spans


class DDPGfD(object):
    def __init__(self, state_dim, action_dim, max_action, n, discount=0.995, tau=0.0005, batch_size=64, device=None,
                 process_group=None, capturable=False):
        self.device = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.actor = Actor(state_dim, action_dim, max_action).to(self.device)
        self.critic = Critic(state_dim, action_dim).to(self.device)
        self._seat()
        self.actor_target = copy.deepcopy(self.actor)
        self.critic_target = copy.deepcopy(self.critic)
        # capturable: Adam keeps its step count on the device so that an update can be captured into a CUDA graph (train.py)
        self.actor_optimizer = torch.optim.Adam(self.actor.parameters(), lr=1e-4, capturable=capturable)
        self.critic_optimizer = torch.optim.Adam(self.critic.parameters(), weight_decay=1e-4, capturable=capturable)
        self.discount = discount
        self.tau = tau
        self.n = n
        self.network_repl_freq = 10
        self.total_it = 0
        self.batch_size = batch_size
        # data-parallel training: None = the default group if torch.distributed is initialised, False = never
        self.process_group = None if process_group in (None, False) else process_group
        self.world = 1
        if process_group is not False and torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(self.process_group)

    def _seat(self):
        # one contiguous bucket for all 310 105 parameters and one for their gradients; the critic's and the
        # actor's gradients are its two halves.  An update needs them at two different times -- the actor loss is taken
        # through the critic AFTER its step (DDPG:345-350) -- so each half is all-reduced when it is complete.
        self.flat_param, self.flat_grad, ((co, cn), (ao, an)) = flatten_parameters([self.critic, self.actor])
        self.critic_flat, self.critic_grad = self.flat_param[co:co + cn], self.flat_grad[co:co + cn]
        self.actor_flat, self.actor_grad = self.flat_param[ao:ao + an], self.flat_grad[ao:ao + an]

    # ------------------------------------------------------------------ acting
    def select_action(self, state):
        """DDPG:71-73 (single observation in, numpy action out)."""
        state = torch.as_tensor(np.asarray(state, dtype=np.float32).reshape(1, -1)).to(self.device)
        with torch.no_grad():
            return self.actor(state).cpu().data.numpy().flatten()

    @torch.no_grad()
    def select_actions(self, states, expl_noise=0.0, max_action=None, generator=None):
        """Batched policy for the vectorised env: [N,82] device tensor -> [N,4]; Gaussian exploration noise and the clip
        to [0, max_action] of main_DDPGfD.py:443-446 stay on the device."""
        a = self.actor(states)
        if expl_noise > 0.0:
            hi = self.actor.max_action if max_action is None else max_action
            a = (a + torch.randn(a.shape, device=a.device, generator=generator) * (hi * expl_noise)).clamp(0.0, hi)
        return a

    # ------------------------------------------------------------------ learning
    def _allreduce(self, flat_grad):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(flat_grad, group=self.process_group)
            flat_grad.div_(self.world)

    def train_on_batch(self, state, action, next_state, reward, not_done=None):
        """One update on explicit n-step windows: state/next_state [B,n,82], action [B,n,4], reward [B,n].
        Returns the four losses as device tensors (actor, critic, critic_L1, critic_LN).  DDPG:263-367."""
        self.total_it += 1
        reward = reward.unsqueeze(-1)
        with torch.no_grad():
            target_Q = self.critic_target(next_state[:, 0], self.actor_target(next_state[:, 0]))
            target_Q = reward[:, 0] + self.discount * target_Q
            target_action = self.actor_target(next_state[:, -1])
            target_critic_val = self.critic_target(next_state[:, -1], target_action)
            n_step_return = torch.zeros(reward.shape[0], device=reward.device)
            for i in range(self.n):       # same accumulation order as DDPG:305-306
                n_step_return += (self.discount ** i) * reward[:, i].squeeze(-1)
            target_QN = (n_step_return + (self.discount ** self.n) * target_critic_val.squeeze(-1)).unsqueeze(-1)
        current_Q = self.critic(state[:, 0], action[:, 0])
        critic_L1loss = F.mse_loss(current_Q, target_Q)
        critic_LNloss = F.mse_loss(current_Q, target_QN)
        critic_loss = critic_L1loss + 0.5 * critic_LNloss
        self.critic_grad.zero_()
        critic_loss.backward()
        self._allreduce(self.critic_grad)
        self.critic_optimizer.step()

        actor_loss = -self.critic(state, self.actor(state)).mean()
        self.actor_grad.zero_()
        actor_loss.backward()          # also deposits gradients in the critic's buffer; they are zeroed before its next use
        self._allreduce(self.actor_grad)
        self.actor_optimizer.step()

        if self.total_it % self.network_repl_freq == 0:
            with torch.no_grad():
                for param, target_param in zip(self.critic.parameters(), self.critic_target.parameters()):
                    target_param.data.copy_(self.tau * param.data + (1 - self.tau) * target_param.data)
                for param, target_param in zip(self.actor.parameters(), self.actor_target.parameters()):
                    target_param.data.copy_(self.tau * param.data + (1 - self.tau) * target_param.data)
        return actor_loss.detach(), critic_loss.detach(), critic_L1loss.detach(), critic_LNloss.detach()

    def train_batch(self, episode_step, expert_replay_buffer, replay_buffer, num_trajectories, prob=0.3):
        """DDPG:219-367, same sampling split (agent batch first, then expert) and same return type (four floats)."""
        if replay_buffer is not None and expert_replay_buffer is None:
            state, action, next_state, reward, not_done = replay_buffer.sample_batch_nstep(self.batch_size, num_trajectories)
        elif replay_buffer is None and expert_replay_buffer is not None:
            state, action, next_state, reward, not_done = expert_replay_buffer.sample_batch_nstep(self.batch_size, num_trajectories)
        else:
            agent_batch_size = int(self.batch_size * (1 - prob))
            expert_batch_size = self.batch_size - agent_batch_size
            a = replay_buffer.sample_batch_nstep(agent_batch_size, num_trajectories)
            e = expert_replay_buffer.sample_batch_nstep(expert_batch_size, num_trajectories)
            state, action, next_state, reward, not_done = (torch.cat((x, y), 0) for x, y in zip(a, e))
        dev = self.device
        losses = self.train_on_batch(state.to(dev), action.to(dev), next_state.to(dev), reward.to(dev), not_done.to(dev))
        return tuple(float(x) for x in torch.stack(losses).tolist())

    # ------------------------------------------------------------------ checkpoint (DDPG:371-382)
    def save(self, filename):
        torch.save(self.critic.state_dict(), filename + "_critic")
        torch.save(self.critic_optimizer.state_dict(), filename + "_critic_optimizer")
        torch.save(self.actor.state_dict(), filename + "_actor")
        torch.save(self.actor_optimizer.state_dict(), filename + "_actor_optimizer")

    def load(self, filename):
        self.critic.load_state_dict(torch.load(filename + "_critic", map_location=self.device))
        self.critic_optimizer.load_state_dict(torch.load(filename + "_critic_optimizer", map_location=self.device))
        self.actor.load_state_dict(torch.load(filename + "_actor", map_location=self.device))
        self.actor_optimizer.load_state_dict(torch.load(filename + "_actor_optimizer", map_location=self.device))
