"""The optimizer tail of a PPO epoch on the agent's flat buffers.

Reference: `compute_gradient_norm(ac.parameters())` (tools/util.py:61-69: one torch.norm launch per parameter tensor),
`torch.nn.utils.clip_grad_norm_` (ppo.py:144) and `Adam(parameters, lr, amsgrad)` (tools/util.py:197-205, ppo.py:145).
`FlatAdam` IS a `torch.optim.Adam` (same constructor arguments after the agent, same `param_groups`, `state_dict()` /
`load_state_dict()` interchangeable with one built on `agent.parameters()`), whose `step()` runs two kernels on the flat
parameter / gradient / moment buffers instead of a multi-tensor sweep over ~100 views:

    optimizer = FlatAdam(agent, lr=3e-4, amsgrad=False)
    ...
    norm = optimizer.step(max_grad_norm=0.5)     # gradient norm + clipping + update; returns the norm as a 0-d device tensor
    optimizer.step()                             # plain torch semantics (clip beforehand with clip_grad_norm_ if wanted)

Data-parallel agents: the pending gradient all-reduce (agents/_flat.py::sync_grads) is flushed first."""
import ctypes

import torch

from molgym_b200 import _cabi


class FlatAdam(torch.optim.Adam):
    def __init__(self, agent, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, maximize=False):
        self._agent = agent
        params = [agent._views[n] for n in agent._p_names]
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad, maximize=maximize,
                         foreach=False, capturable=False, differentiable=False, fused=False)
        self._bind_state()

    def _bind_state(self):
        """Moments as flat buffers; the per-parameter state tensors torch's state_dict() walks are views of them."""
        agent = self._agent
        dev, n = agent._flat.device, agent._flat.numel()
        lib = agent._rt.lib()
        self._flat_ptr = agent._flat.data_ptr()
        self._m = torch.zeros(n, dtype=torch.float32, device=dev)
        self._v = torch.zeros(n, dtype=torch.float32, device=dev)
        self._vmax = torch.zeros(n, dtype=torch.float32, device=dev) if self.param_groups[0]['amsgrad'] else None
        self._scratch = torch.zeros(lib.mgb_optim_scratch_bytes(), dtype=torch.uint8, device=dev)
        self._norm = torch.zeros(2, dtype=torch.float64, device=dev)
        self._steps = 0
        self._step_t = torch.tensor(0.0, dtype=torch.float32)   # one host counter shared by every parameter's state entry
        for p, o, k in zip(agent._param_list, agent._p_offsets, agent._p_numels):
            old = self.state.get(p, {})
            st = {'step': self._step_t,
                  'exp_avg': self._m[o:o + k].view(p.shape), 'exp_avg_sq': self._v[o:o + k].view(p.shape)}
            if self._vmax is not None:
                st['max_exp_avg_sq'] = self._vmax[o:o + k].view(p.shape)
            for key in ('exp_avg', 'exp_avg_sq', 'max_exp_avg_sq'):
                if key in old and key in st:
                    st[key].copy_(old[key])
            if 'step' in old:
                self._steps = int(float(old['step']))
                self._step_t.fill_(float(self._steps))
            self.state[p] = st

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)   # replaces the state tensors by copies of the loaded ones ...
        self.param_groups[0]['foreach'] = False
        self._bind_state()                    # ... which are folded back into the flat buffers

    def zero_grad(self, set_to_none: bool = True) -> None:
        """One memset of the flat gradient buffer; the parameters' `.grad` views stay attached (agents/_flat.py::zero_grad_flat)."""
        self._agent.zero_grad_flat()

    def grad_norm(self) -> torch.Tensor:
        """||grad||_2 over all parameters as a 0-d float64 device tensor (compute_gradient_norm, tools/util.py:61-69)."""
        agent = self._agent
        agent.sync_grads()
        agent._attach_grads()
        lib, rt = agent._rt.lib(), agent._rt
        with rt.device_ctx():
            _cabi.check(lib, lib.mgb_grad_norm(agent._flat_grad.data_ptr(), agent._flat_grad.numel(), self._scratch.data_ptr(),
                                               self._norm.data_ptr(), rt.stream_ptr()))
        return self._norm[0]

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm=None, reuse_norm=False):
        """One Adam / AMSGrad update of every parameter.  With `max_grad_norm` the gradient norm is computed (or, `reuse_norm`,
        taken from the grad_norm() call just made) and the gradients are clipped inside the update (clip_grad_norm_ semantics) —
        the norm is returned as a 0-d device tensor."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        agent = self._agent
        if not agent._params_aliased() or agent._flat.data_ptr() != self._flat_ptr:
            raise RuntimeError('FlatAdam: the agent\'s parameters moved (unpickled / .to()); build a new optimizer for it')
        if getattr(agent, '_grads_fresh', False) or all(p.grad is None for p in agent._param_list):
            return loss        # nothing was differentiated since zero_grad(): torch skips parameters without gradients
        norm = None
        if max_grad_norm is not None:
            norm = self._norm[0] if reuse_norm else self.grad_norm()       # also flushes a pending gradient all-reduce
        agent.sync_grads()
        agent._attach_grads()
        g = self.param_groups[0]
        self._steps += 1
        lib, rt = agent._rt.lib(), agent._rt
        with rt.device_ctx():
            _cabi.check(lib, lib.mgb_adam_step(agent._flat.data_ptr(), agent._flat_grad.data_ptr(), self._m.data_ptr(), self._v.data_ptr(),
                                               self._vmax.data_ptr() if self._vmax is not None else None, agent._flat.numel(),
                                               float(g['lr']), float(g['betas'][0]), float(g['betas'][1]), float(g['eps']),
                                               float(g['weight_decay']), self._steps, 1 if g['amsgrad'] else 0,
                                               1 if g['maximize'] else 0, self._norm.data_ptr() if norm is not None else None,
                                               float(max_grad_norm) if max_grad_norm is not None else 0.0, rt.stream_ptr()))
        torch.autograd.graph.increment_version(agent._flat)   # the kernel wrote the buffer behind autograd's back: the fused step
        self._step_t += 1                                      # watches this counter to order itself behind parameter updates
        return norm if max_grad_norm is not None else loss
