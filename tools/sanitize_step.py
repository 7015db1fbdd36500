"""Target of compute-sanitizer (memcheck / racecheck / synccheck): two PPO minibatch steps of the covariant agent through the fused
CUDA-graph-free path (eager launches: the sanitizer instruments kernels launched directly), one evaluate-mode step + backward, and
one FlatAdam step, rollouts (sample + greedy), and one step of the internal-coordinate (SchNet) agent.  MGB_MIX_TC=1 routes the channel mix
through the tensor-core kernels.
    compute-sanitizer --tool memcheck python tools/sanitize_step.py [workload] [batch]"""
import dataclasses
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from molgym_b200 import ppo, synth  # noqa: E402
from molgym_b200.agents.covariant.agent import CovariantAC  # noqa: E402
from molgym_b200.agents.internal.agent import SchNetAC  # noqa: E402
from molgym_b200.spaces import ActionSpace, ObservationSpace  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'C2'
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 24
cfg = synth.CONFIGS[which]
dev = torch.device('cuda:0')
torch.manual_seed(0)
agent = CovariantAC(ObservationSpace(cfg.canvas_size, cfg.zs), ActionSpace(cfg.zs), device=dev, **cfg.agent_kwargs())
agent.graph_evaluate = False      # eager launches
agent.fused_ppo = False
obs, n = synth.make_observations(cfg, batch=batch)
act = synth.make_actions(cfg, obs, n)
with torch.no_grad():
    logp0 = agent.step(obs, act)['logp'].cpu().numpy()
old_logp, adv, ret = synth.make_ppo_targets(cfg, logp0)
data = dict(obs=obs, act=act, logp=old_logp, adv=adv, ret=ret)
for _ in range(2):
    agent.zero_grad()
    loss, info = ppo.compute_loss(agent, data, 0.2, 0.5, 0.01)
    loss.backward()
from molgym_b200.optim import FlatAdam  # noqa: E402
opt = FlatAdam(agent, lr=3e-4)
opt.grad_norm()
opt.step(max_grad_norm=0.5, reuse_norm=True)   # k_grad_norm + k_adam_step
torch.cuda.synchronize()
print('covariant', which, batch, 'loss', float(loss), 'grad norm', float(torch.cat([p.grad.reshape(-1) for p in agent.parameters()]).norm()))
for training in (True, False):
    agent.training = training
    with torch.no_grad():
        pred = agent.step(obs[:6])
torch.cuda.synchronize()
print('rollout ok', pred['a'].shape)
c1 = synth.CONFIGS['C1']
sch = SchNetAC(ObservationSpace(c1.canvas_size, c1.zs), ActionSpace(c1.zs), device=dev, **c1.agent_kwargs())
o1, n1 = synth.make_observations(c1, batch=12)
a1 = synth.make_actions(c1, o1, n1)
p = sch.step(o1, a1)
(p['logp'].sum() + p['v'].sum() + p['ent'].sum()).backward()
torch.cuda.synchronize()
print('internal ok', float(p['logp'].sum()))
