"""Flat parameter / gradient storage shared by the agents: every nn.Parameter is a view of one fp32 buffer laid out as
the C ABI expects, every `.grad` a view of one gradient buffer the CUDA backward writes into (and the one NCCL
all-reduce of the data-parallel path runs over)."""
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


class FlatParamMixin:
    """Needs: self.device, self._p_names / _p_offsets / _p_numels / _p_total (from the plan)."""

    def _init_flat(self, shapes: Dict[str, tuple], values: Dict[str, torch.Tensor], order: List[str]):
        self._flat = torch.zeros(self._p_total, dtype=torch.float32, device=self.device)
        self._flat_grad = torch.zeros(self._p_total, dtype=torch.float32, device=self.device)
        self._grad_scratch = None
        index = {n: i for i, n in enumerate(self._p_names)}
        host = torch.zeros(self._p_total, dtype=torch.float32)
        for name in self._p_names:
            i = index[name]
            o, n = self._p_offsets[i], self._p_numels[i]
            assert int(np.prod(shapes[name])) == n, (name, shapes[name], n)
            host[o:o + n] = values[name].reshape(-1).to(torch.float32)
        self._flat.copy_(host)
        self._views, self._grad_views = {}, {}
        for name in order:
            i = index[name]
            o, n = self._p_offsets[i], self._p_numels[i]
            p = nn.Parameter(self._flat[o:o + n].view(shapes[name]), requires_grad=True)
            register_dotted(self, name, p)
            self._views[name] = p
            self._grad_views[name] = self._flat_grad[o:o + n].view(shapes[name])
        self._param_list = [self._views[n] for n in self._p_names]

    def _rebuild_flat_after_unpickle(self):
        named = dict(self.named_parameters())
        self._flat = torch.zeros(self._p_total, dtype=torch.float32, device=self.device)
        self._flat_grad = torch.zeros(self._p_total, dtype=torch.float32, device=self.device)
        self._grad_scratch = None
        self._views = {n: named[n] for n in self._p_names}
        self._param_list = [self._views[n] for n in self._p_names]
        self._grad_views = {n: self._flat_grad[o:o + k].view(self._views[n].shape)
                            for n, o, k in zip(self._p_names, self._p_offsets, self._p_numels)}
        self._realias()

    # keep nn.Parameters aliased to the flat buffers (load_state_dict / optimizers keep the aliasing; .to(), pickling or
    # user code that rebinds .data do not)
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
                    p.data = view

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
        if set_to_none:
            for p in self._param_list:
                p.grad = None
        else:
            self._flat_grad.zero_()
            for p in self._param_list:
                if p.grad is not None and p.grad.data_ptr() != self._flat_grad.data_ptr():
                    p.grad.zero_()

    def _is_sharded(self) -> bool:
        return bool(getattr(self, 'data_parallel', False)) and torch.distributed.is_available() and \
            torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1

    def _grad_target(self, keep: bool):
        """(tensor the CUDA backward writes into, accumulate flag)."""
        if self._is_sharded():
            if self._grad_scratch is None:
                self._grad_scratch = torch.empty_like(self._flat_grad)
            return self._grad_scratch, 0
        return self._flat_grad, 1 if keep else 0

    def _finish_grads(self, keep: bool):
        if self._is_sharded():
            # the one exchange step of the path: sum of the flat gradient over ranks (NCCL over NVLink)
            torch.distributed.all_reduce(self._grad_scratch, op=torch.distributed.ReduceOp.SUM)
            if keep:
                self._flat_grad.add_(self._grad_scratch)
            else:
                self._flat_grad.copy_(self._grad_scratch)
