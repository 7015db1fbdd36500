"""Experiment: one minibatch as two concurrent half-batch chains (two graphs on two streams) vs one chain."""
import ctypes, os, sys, time
import torch
sys.path.insert(0, '/root/repo')
from molgym_b200 import _cabi, _lib, synth
from molgym_b200.agents.covariant.agent import CovariantAC
from molgym_b200.spaces import ActionSpace, ObservationSpace

cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'C2']
B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg.mini_batch_size
parts = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device('cuda:0')
lib = _lib.load()
torch.manual_seed(0)
agent = CovariantAC(ObservationSpace(cfg.canvas_size, cfg.zs), ActionSpace(cfg.zs), device=dev, **cfg.agent_kwargs())
obs, n = synth.make_observations(cfg, batch=B)
act = synth.make_actions(cfg, obs, n)
parsed = agent.parse_observations(obs)
pos, charges, bags = parsed['positions'].contiguous(), parsed['charges'].contiguous(), parsed['bags'].contiguous()
act_d = torch.as_tensor(act, dtype=torch.float32, device=dev)
f32 = dict(dtype=torch.float32, device=dev)
old = torch.zeros(B, **f32); adv = torch.randn(B, dtype=torch.float64, device=dev); ret = torch.randn(B, dtype=torch.float64, device=dev)

def make(lo, hi):
    b = hi - lo
    out = torch.empty(6, b, **f32)
    info = torch.zeros(8, dtype=torch.float64, device=dev)
    grad = torch.zeros_like(agent._flat)
    ws = torch.empty(lib.mgb_cov_workspace_bytes(agent._plan, b), dtype=torch.uint8, device=dev)
    o = _cabi.CovOutputs(); o.logp, o.ent, o.v = out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr()
    p_, c_, g_, a_ = pos[lo:hi], charges[lo:hi], bags[lo:hi], act_d[lo:hi]
    def run(stream):
        _cabi.check(lib, lib.mgb_cov_forward(agent._plan, b, p_.data_ptr(), c_.data_ptr(), g_.data_ptr(), a_.data_ptr(), agent._flat.data_ptr(), ws.data_ptr(), ws.numel(), ctypes.byref(o), stream))
        _cabi.check(lib, lib.mgb_ppo_loss(b, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), old[lo:hi].data_ptr(), adv[lo:hi].data_ptr(), ret[lo:hi].data_ptr(), 0.2, 0.5, 0.01, 1.0 / B, info.data_ptr(), out[3].data_ptr(), out[4].data_ptr(), out[5].data_ptr(), stream))
        _cabi.check(lib, lib.mgb_cov_backward(agent._plan, b, p_.data_ptr(), c_.data_ptr(), g_.data_ptr(), a_.data_ptr(), agent._flat.data_ptr(), ws.data_ptr(), ws.numel(), out[3].data_ptr(), out[4].data_ptr(), out[5].data_ptr(), grad.data_ptr(), 0, stream))
    run(torch.cuda.current_stream(dev).cuda_stream); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); cs = torch.cuda.Stream(dev, priority=-5); cs.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(cs):
        with torch.cuda.graph(g, stream=cs):
            run(torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.current_stream(dev).wait_stream(cs)
    return g, (out, info, grad, ws, o, p_, c_, g_, a_)

bounds = [(B * k // parts, B * (k + 1) // parts) for k in range(parts)]
graphs = [make(lo, hi) for lo, hi in bounds]
streams = [torch.cuda.Stream(dev) for _ in range(parts)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def step():
    cur = torch.cuda.current_stream(dev)
    for (g, _), st in zip(graphs, streams):
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            g.replay()
    for st in streams:
        cur.wait_stream(st)

for _ in range(10): step()
torch.cuda.synchronize()
K = 100
tot = 0.0
for _ in range(K):
    flush.zero_()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); step(); b_.record()
    torch.cuda.synchronize()
    tot += a.elapsed_time(b_)
print(cfg.name, 'B', B, 'parts', parts, 'ms/step', tot / K)
