"""Host-side mirror of the PPO minibatch step that calls the hot path (reference: molgym/ppo.py:18-63 compute_loss,
:66-89 batch generation, :99-160 train).  The reference's own ppo.py runs unchanged on top of molgym_b200's agents
(tests/test_dropin_reference_loop.py); this restatement exists so the benchmark and the tests can drive the same arithmetic
on machines where the reference tree is not present (the GPU box), and so that `train` can take the fused step and the fused
optimizer tail (molgym_b200.optim.FlatAdam) when the caller hands them over."""
import time
from typing import Optional, Dict, Tuple

import numpy as np
import torch


def _to_device(x, device):
    """torch.as_tensor(x, device=device) (ppo.py:28-30) without stalling the host behind the forward pass that was just
    enqueued: numpy arrays go through a pinned staging copy and a non-blocking transfer."""
    if isinstance(x, np.ndarray) and torch.device(device).type == 'cuda':
        return torch.from_numpy(x).pin_memory().to(device, non_blocking=True)
    return torch.as_tensor(x, device=device)


def _takes_fused_step(ac, data: dict, device) -> bool:
    """The agent offers the fused step and the buffers have the dtypes ppo.train hands over (buffer.py:106-116)."""
    same_device = device is None or torch.device(device) == getattr(ac, 'device', None) or \
        (torch.device(device).type == 'cuda' and torch.device(device).index is None and getattr(ac, 'device', torch.device('cpu')).type == 'cuda')
    return bool(getattr(ac, 'fused_ppo', False) and same_device and torch.is_grad_enabled() and isinstance(data['adv'], np.ndarray)
                and data['adv'].dtype == np.float64 and isinstance(data['ret'], np.ndarray) and data['ret'].dtype == np.float64
                and isinstance(data['logp'], np.ndarray) and data['logp'].dtype == np.float32 and len(data['obs']) > 0)


def compute_loss(ac, data: dict, clip_ratio: float, vf_coef: float, entropy_coef: float, device=None,
                 n_global: Optional[int] = None) -> Tuple[torch.Tensor, Dict[str, float]]:
    """ppo.py:18-63, same operations in the same order (adv / ret arrive as float64 numpy -> float64 loss).

    Agents that offer `fused_ppo_loss` (CovariantAC) take the whole step — packing, one H2D copy, forward, the same loss
    arithmetic in float64 on the device (k_ppo_loss, checked against this function by the tests) and the backward — as CUDA-graph
    replays when the buffers have the dtypes ppo.train hands over (buffer.py:106-116); set `ac.fused_ppo = False` for the
    op-by-op path below.  `n_global` (data-parallel agents, fused step only): `data` is already this rank's slice
    (`ac.train_shard`) of a minibatch of n_global canvases."""
    if _takes_fused_step(ac, data, device):
        if n_global is not None:
            return ac.fused_ppo_loss(data['obs'], data['act'], data['logp'], data['adv'], data['ret'], clip_ratio, vf_coef, entropy_coef,
                                     n_global=n_global)
        return ac.fused_ppo_loss(data['obs'], data['act'], data['logp'], data['adv'], data['ret'], clip_ratio, vf_coef, entropy_coef)
    if n_global is not None:
        raise ValueError('compute_loss: n_global (a pre-sliced minibatch) needs the fused step')
    pred = ac.step(data['obs'], data['act'])
    device = device if device is not None else pred['logp'].device
    old_logp = _to_device(data['logp'], device)
    adv = _to_device(data['adv'], device)
    ret = _to_device(data['ret'], device)
    ratio = torch.exp(pred['logp'] - old_logp)
    obj = ratio * adv
    clipped_obj = ratio.clamp(1 - clip_ratio, 1 + clip_ratio) * adv
    policy_loss = -torch.min(obj, clipped_obj).mean()
    entropy_loss = -entropy_coef * pred['ent'].mean()
    vf_loss = vf_coef * (pred['v'] - ret).pow(2).mean()
    loss = policy_loss + entropy_loss + vf_loss
    approx_kl = (old_logp - pred['logp']).mean()
    clipped = ratio.lt(1 - clip_ratio) | ratio.gt(1 + clip_ratio)
    clip_fraction = torch.as_tensor(clipped, dtype=torch.float32).mean()
    stats = torch.stack([policy_loss.detach().double(), entropy_loss.detach().double(), vf_loss.detach().double(),
                         loss.detach().double(), approx_kl.detach().double(), clip_fraction.double()]).cpu().numpy()
    info = dict(policy_loss=float(stats[0]), entropy_loss=float(stats[1]), vf_loss=float(stats[2]), total_loss=float(stats[3]),
                approx_kl=float(stats[4]), clip_fraction=float(stats[5]))
    return loss, info


def get_batch_generator(indices: np.ndarray, batch_size: int):
    """ppo.py:66-74."""
    assert len(indices.shape) == 1
    indices = np.random.permutation(indices)
    batches = indices[:len(indices) // batch_size * batch_size].reshape(-1, batch_size)
    for batch in batches:
        yield batch
    r = len(indices) % batch_size
    if r:
        yield indices[-r:]


def collect_data_batch(data: dict, indices: np.ndarray) -> dict:
    """ppo.py:77-89."""
    batch = {}
    for k, v in data.items():
        batch[k] = [v[i] for i in indices] if isinstance(v, list) else v[indices]
    return batch


def compute_gradient_norm(parameters, norm_type: int = 2) -> float:
    """tools/util.py:61-69."""
    parameters = [p for p in parameters if p.grad is not None]
    if len(parameters) == 0:
        return 0.0
    device = parameters[0].grad.device
    return torch.norm(torch.stack([torch.norm(p.grad.detach(), norm_type).to(device) for p in parameters]), norm_type).item()


def train(ac, optimizer, data: dict, mini_batch_size: int, clip_ratio: float, target_kl: float, vf_coef: float, entropy_coef: float,
          gradient_clip: float, max_num_steps: int, device=None) -> dict:
    """ppo.py:99-160, same control flow and the same returned info.  With a `molgym_b200.optim.FlatAdam` optimizer the gradient
    norm, the clipping and the Adam update run as two kernels on the flat buffers; any other optimizer takes the reference's
    sequence (compute_gradient_norm, clip_grad_norm_, optimizer.step)."""
    from molgym_b200.optim import FlatAdam
    infos: Dict[str, float] = {}
    start_time = time.time()
    num_epochs = 0
    flat = isinstance(optimizer, FlatAdam)
    kept_norm = None   # FlatAdam: the gradient norm of the last completed step stays on the device until the loop is over
    for i in range(max_num_steps):
        optimizer.zero_grad()
        batch_infos = []
        for batch_indices in get_batch_generator(indices=np.arange(len(data['obs'])), batch_size=mini_batch_size):
            # data-parallel agent on the fused step: collect only this rank's slice of the minibatch (every rank draws the same
            # permutation), so that the host work per rank does not grow with the number of ranks
            shard = ac.train_shard(len(batch_indices)) if (hasattr(ac, 'train_shard') and _takes_fused_step(ac, data, device)) else None
            if shard is not None:
                data_batch = collect_data_batch(data, indices=batch_indices[shard[0]:shard[1]])
                batch_loss, batch_info = compute_loss(ac, data=data_batch, clip_ratio=clip_ratio, vf_coef=vf_coef, entropy_coef=entropy_coef,
                                                      device=device, n_global=len(batch_indices))
            else:
                data_batch = collect_data_batch(data, indices=batch_indices)
                batch_loss, batch_info = compute_loss(ac, data=data_batch, clip_ratio=clip_ratio, vf_coef=vf_coef, entropy_coef=entropy_coef,
                                                      device=device)
            batch_loss.backward(retain_graph=False)
            batch_infos.append(batch_info)
        # the loss numbers come from the forward passes: reading them does not wait for the backward of the last minibatch
        loss_info = {key: np.mean([d[key] for d in batch_infos]) for key in batch_infos[0].keys()}
        if flat:
            norm = optimizer.grad_norm().clone()   # only the infos of the last completed step are returned (ppo.py:150): no read-back here
        else:
            loss_info['grad_norm'] = compute_gradient_norm(ac.parameters())
        if loss_info['approx_kl'] > 1.5 * target_kl:
            break
        if flat:
            optimizer.step(max_grad_norm=gradient_clip, reuse_norm=True)
            kept_norm = norm
        else:
            torch.nn.utils.clip_grad_norm_(ac.parameters(), max_norm=gradient_clip)
            optimizer.step()
        optimizer.zero_grad()
        num_epochs += 1
        infos.update(loss_info)
    if kept_norm is not None:
        infos['grad_norm'] = float(kept_norm.item())
    infos['num_opt_steps'] = num_epochs
    infos['time'] = time.time() - start_time
    return infos
