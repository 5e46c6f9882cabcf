"""CPU port of the reference's Rainbow update (BENCH INFRASTRUCTURE ONLY:
the timed "reference CPU path" of bench.py's Rainbow section, kind="port").

Restates with plain torch on the host, single process, torch intra-op threads:
  pfrl/q_functions/dueling_dqn.py:67-129   DistributionalDuelingDQN
  pfrl/nn/noisy_linear.py:25-70            FactorizedNoisyLinear
  pfrl/agents/categorical_double_dqn.py:10-52, categorical_dqn.py:7-57,178-204
  pfrl/agents/dqn.py:316-365               update(): batch_experiences ->
      loss -> update_errors -> backward -> Adam step
on top of oracle/pyport.py's replay buffer.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle.pyport import py_batch_experiences


class NoisyLinear(nn.Module):
    def __init__(self, n_in, n_out, sigma_scale=0.5):
        super().__init__()
        self.mu = nn.Linear(n_in, n_out)
        self.sigma = nn.Linear(n_in, n_out)
        with torch.no_grad():
            self.mu.weight.uniform_(-1 / math.sqrt(n_in), 1 / math.sqrt(n_in))
            self.sigma.weight.fill_(sigma_scale / math.sqrt(n_in))
            self.sigma.bias.fill_(sigma_scale / math.sqrt(n_out))

    def forward(self, x):
        n_out, n_in = self.sigma.weight.shape
        r = torch.randn(n_in + n_out)
        eps = torch.sign(r) * torch.sqrt(torch.abs(r))
        w = torch.addcmul(self.mu.weight, self.sigma.weight, torch.outer(eps[n_in:], eps[:n_in]))
        b = torch.addcmul(self.mu.bias, self.sigma.bias, eps[n_in:])
        return F.linear(x, w, b)


class RainbowNet(nn.Module):
    def __init__(self, n_actions=18, n_atoms=51, v_min=-10, v_max=10):
        super().__init__()
        self.n_actions, self.n_atoms = n_actions, n_atoms
        self.z = torch.linspace(v_min, v_max, n_atoms)
        self.convs = nn.ModuleList([nn.Conv2d(4, 32, 8, stride=4), nn.Conv2d(32, 64, 4, stride=2),
                                    nn.Conv2d(64, 64, 3, stride=1)])
        self.main = NoisyLinear(3136, 1024)
        self.adv = NoisyLinear(512, n_actions * n_atoms)
        self.val = NoisyLinear(512, n_atoms)

    def forward(self, x):
        h = x
        for c in self.convs:
            h = torch.relu(c(h))
        h = torch.relu(self.main(h.reshape(x.shape[0], -1)))
        ha, hv = torch.chunk(h, 2, dim=1)
        a = self.adv(ha).reshape(-1, self.n_actions, self.n_atoms)
        a = a - a.sum(1, keepdim=True) / self.n_actions
        return F.softmax(a + self.val(hv).reshape(-1, 1, self.n_atoms), dim=2)


def projection(Tz, probs, z):
    n = z.shape[0]
    dz = z[1] - z[0]
    bj = torch.clamp((torch.clamp(Tz, z[0], z[-1]) - z[0]) / dz, 0, n - 1)
    lo, up = torch.floor(bj), torch.ceil(bj)
    out = torch.zeros_like(probs)
    out.scatter_add_(1, lo.long(), probs * (1 - (bj - lo)))
    out.scatter_add_(1, up.long(), probs * (bj - lo))
    return out


class PyRainbow:
    def __init__(self, rbuf, gamma=0.99, batch=512, n_actions=18):
        self.model = RainbowNet(n_actions)
        self.target = RainbowNet(n_actions)
        self.target.load_state_dict(self.model.state_dict())
        self.opt = torch.optim.Adam(self.model.parameters(), 6.25e-5, eps=1.5e-4)
        self.rbuf, self.gamma, self.batch = rbuf, gamma, batch
        self.phi = lambda x: np.asarray(x, dtype=np.float32) / 255

    def act(self, obs_list):
        with torch.no_grad():
            x = torch.as_tensor(np.stack([self.phi(o) for o in obs_list]))
            q = (self.model(x) * self.model.z).sum(2)
            return q.argmax(1).numpy()

    def update(self):
        exps = self.rbuf.sample(self.batch)
        b = py_batch_experiences(exps, torch.device("cpu"), self.phi, self.gamma)
        w = torch.tensor([e[0]["weight"] for e in exps], dtype=torch.float32)
        n = self.batch
        ar = torch.arange(n)
        y = self.model(b["state"])[ar, b["action"].long()]
        with torch.no_grad():
            z = self.model.z
            nxt_t = self.target(b["next_state"])
            nxt_o = self.model(b["next_state"])
            a_star = (nxt_o * z).sum(2).argmax(1)
            Tz = b["reward"][:, None] + (1 - b["is_state_terminal"][:, None]) * \
                b["discount"][:, None] * z[None]
            t = projection(Tz, nxt_t[ar, a_star], z)
        elt = -t * torch.log(torch.clamp(y, 1e-10, 1.0))
        per = elt.sum(1)
        loss = torch.matmul(per, w) / n
        self.rbuf.update_errors(per.detach().numpy())
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(loss.detach())
