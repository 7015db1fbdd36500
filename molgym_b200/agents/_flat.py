"""Flat parameter / gradient storage shared by the agents: every nn.Parameter is a view of one fp32 buffer laid out as
the C ABI expects, every `.grad` a view of one gradient buffer the CUDA backward writes into.

Data-parallel agents (molgym_b200.parallel.shard_agent) accumulate the gradients of their shard in a second flat buffer;
the path's one exchange step — a sum-all-reduce of that buffer (NCCL over NVLink) — runs once per optimizer step, when
the gradients are first read (`parameters()`, which ppo.train calls at molgym/ppo.py:135 before clipping and stepping, or
any optimizer's step()), not once per minibatch: ppo.train accumulates over the minibatches of an epoch first
(molgym/ppo.py:122-131)."""
import weakref
from typing import Dict, List

import numpy as np
import torch
import torch.nn as nn


class _Node(nn.Module):
    """Anonymous container used to reproduce the reference's dotted parameter names."""


def register_dotted(root: nn.Module, dotted: str, param: nn.Parameter):
    parts = dotted.split('.')
    node = root
    for part in parts[:-1]:
        if part not in node._modules:
            node.add_module(part, _Node())
        node = node._modules[part]
    node.register_parameter(parts[-1], param)


_SHARDED_AGENTS = weakref.WeakSet()
_HOOK = []


def _optimizer_pre_hook(optimizer, args, kwargs):
    """Safety net: an optimizer step never sees gradients whose all-reduce is still pending."""
    for agent in list(_SHARDED_AGENTS):
        if agent._grad_pending:
            agent.sync_grads()


def _install_optimizer_hook():
    if not _HOOK:
        from torch.optim.optimizer import register_optimizer_step_pre_hook
        _HOOK.append(register_optimizer_step_pre_hook(_optimizer_pre_hook))


class FlatParamMixin:
    """Needs: self.device, self._p_names / _p_offsets / _p_numels / _p_total (from the plan)."""

    _grad_pending = False

    def _init_flat(self, shapes: Dict[str, tuple], values: Dict[str, torch.Tensor], order: List[str]):
        self._flat = torch.zeros(self._p_total, dtype=torch.float32, device=self.device)
        self._flat_grad = torch.zeros(self._p_total, dtype=torch.float32, device=self.device)
        self._grad_local = None
        self._grad_pending = False
        index = {n: i for i, n in enumerate(self._p_names)}
        host = torch.zeros(self._p_total, dtype=torch.float32)
        for name in self._p_names:
            i = index[name]
            o, n = self._p_offsets[i], self._p_numels[i]
            assert int(np.prod(shapes[name])) == n, (name, shapes[name], n)
            host[o:o + n] = values[name].reshape(-1).to(torch.float32)
        self._flat.copy_(host)
        self._bind_views(shapes, order)

    def _bind_views(self, shapes: Dict[str, tuple], order: List[str]):
        """(Re)create every nn.Parameter as a view of the flat buffer.  Views made this way share the flat buffer's version
        counter, so an in-place update of any parameter (optimizer.step) is visible in `_flat._version`."""
        index = {n: i for i, n in enumerate(self._p_names)}
        self._views, self._grad_views = {}, {}
        for name in order:
            i = index[name]
            o, n = self._p_offsets[i], self._p_numels[i]
            p = nn.Parameter(self._flat[o:o + n].view(shapes[name]), requires_grad=True)
            register_dotted(self, name, p)
            self._views[name] = p
            self._grad_views[name] = self._flat_grad[o:o + n].view(shapes[name])
        self._param_list = [self._views[n] for n in self._p_names]
        self._shared_version = True

    def _rebuild_flat_after_unpickle(self):
        """torch.load / pickle hand back independent parameter tensors: copy them into a new flat buffer and replace them by
        views of it (fresh Parameter objects: nothing else can hold the unpickled ones yet)."""
        named = dict(self.named_parameters())
        order = [n for n in named if n in set(self._p_names)]
        dev = self.device
        self._flat = torch.zeros(self._p_total, dtype=torch.float32, device=dev)
        self._flat_grad = torch.zeros(self._p_total, dtype=torch.float32, device=dev)
        self._grad_local = None
        self._grad_pending = False
        shapes = {}
        with torch.no_grad():
            for name, o, n in zip(self._p_names, self._p_offsets, self._p_numels):
                shapes[name] = tuple(named[name].shape)
                self._flat[o:o + n].copy_(named[name].detach().reshape(-1).to(device=dev, dtype=torch.float32))
        self._bind_views(shapes, order)

    # keep nn.Parameters aliased to the flat buffers (load_state_dict / optimizers keep the aliasing; .to() or user code that
    # rebinds .data do not)
    def _params_aliased(self) -> bool:
        base = self._flat.data_ptr()
        first, last = self._param_list[0], self._param_list[-1]
        return (first.data_ptr() == base + 4 * self._p_offsets[0] and last.data_ptr() == base + 4 * self._p_offsets[-1]
                and first.device == self._flat.device)

    def _realias(self):
        with torch.no_grad():
            for p, o, n in zip(self._param_list, self._p_offsets, self._p_numels):
                view = self._flat[o:o + n].view(p.shape)
                if p.data_ptr() != view.data_ptr():
                    view.copy_(p.data.to(self._flat.device))
                    p.data = view          # gives p its own version counter: fall back to per-parameter versions
                    self._shared_version = False

    def _param_version(self):
        """Changes whenever any parameter was modified in place since the last call site looked."""
        if self._shared_version:
            return (self._flat._version, self._flat.data_ptr())
        return (sum(p._version for p in self._param_list), self._flat.data_ptr())

    def _attach_grads(self) -> bool:
        """Make every p.grad a view of the flat gradient buffer; returns True if existing values must be kept."""
        first = self._param_list[0]
        if first.grad is not None and first.grad.data_ptr() == self._flat_grad.data_ptr() + 4 * self._p_offsets[0]:
            return True
        self._flat_grad.zero_()
        keep = False
        if any(p.grad is not None for p in self._param_list):   # somebody assigned their own gradient tensors: fold them in
            for name, p in zip(self._p_names, self._param_list):
                if p.grad is not None:
                    self._grad_views[name].add_(p.grad)
            keep = True
        for name, p in zip(self._p_names, self._param_list):
            p.grad = self._grad_views[name]
        return keep

    def zero_grad(self, set_to_none: bool = True) -> None:
        """nn.Module.zero_grad walks the module tree (hundreds of tiny containers here); the parameter list is known."""
        self._grad_pending = False   # gradients of the shard that were never reduced are dropped with the rest
        if set_to_none:
            for p in self._param_list:
                p.grad = None
        else:
            self._flat_grad.zero_()
            for p in self._param_list:
                if p.grad is not None and p.grad.data_ptr() != self._flat_grad.data_ptr():
                    p.grad.zero_()

    def zero_grad_flat(self) -> None:
        """zero_grad for callers that own the flat layout (FlatAdam): ONE memset of the flat gradient, the `.grad` views stay
        attached (detaching and re-attaching ~100 views costs ~0.3 ms of host time per optimizer step, right where the next forward
        waits for the host).  `_grads_fresh` records that nothing was accumulated since, so that an optimizer step without a backward
        still skips the update like torch does for `.grad is None`."""
        self._grad_pending = False
        first = self._param_list[0]
        if first.grad is not None and first.grad.data_ptr() == self._flat_grad.data_ptr() + 4 * self._p_offsets[0]:
            self._flat_grad.zero_()
            self._grads_fresh = True
        else:
            self.zero_grad(set_to_none=True)

    # ------------------------------------------------------------------------------------------------------
    # data-parallel gradient exchange
    # ------------------------------------------------------------------------------------------------------
    def _is_sharded(self) -> bool:
        return bool(getattr(self, 'data_parallel', False)) and torch.distributed.is_available() and \
            torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1

    def _grad_target(self, keep: bool):
        """(tensor the backward kernels write into, accumulate flag).  Sharded: the shard-local accumulation buffer, reduced
        later by sync_grads(); else the flat gradient itself."""
        self._grads_fresh = False       # a backward is about to write gradients
        if self._is_sharded():
            if self._grad_local is None:
                self._grad_local = torch.zeros_like(self._flat_grad)
            if not keep:
                self._grad_pending = False     # .grad was (re)created empty: stale shard-local sums go with it
            accumulate = 1 if self._grad_pending else 0
            self._grad_pending = True
            _SHARDED_AGENTS.add(self)
            _install_optimizer_hook()
            return self._grad_local, accumulate
        return self._flat_grad, 1 if keep else 0

    def sync_grads(self):
        """The path's one exchange step: sum the shard-local gradient over ranks and add it to `.grad`.  Called when the
        gradients are first read (parameters()) or before any optimizer step; a no-op when nothing is pending."""
        if not self._grad_pending:
            return
        self._grad_pending = False
        self._before_grad_sync()
        torch.distributed.all_reduce(self._grad_local, op=torch.distributed.ReduceOp.SUM)
        self._flat_grad.add_(self._grad_local)

    def _before_grad_sync(self):
        """Hook: order the caller's stream behind work that still writes the shard-local gradient."""

    def parameters(self, recurse: bool = True):
        self.sync_grads()
        return super().parameters(recurse)
