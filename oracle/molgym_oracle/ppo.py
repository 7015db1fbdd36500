"""ORACLE / TEST INFRASTRUCTURE — restated PPO-clip loss (reference: molgym/ppo.py:18-63)."""
import torch


def ppo_loss(logp, ent, v, old_logp, adv, ret, clip_ratio, vf_coef, entropy_coef):
    """Same arithmetic as ppo.py:28-52.  `adv`/`ret` arrive as float64 numpy in the reference
    (buffer.py:106-114) so the products are promoted to float64; callers pass them in that dtype."""
    old_logp = torch.as_tensor(old_logp, device=logp.device)
    adv = torch.as_tensor(adv, device=logp.device)
    ret = torch.as_tensor(ret, device=logp.device)
    ratio = torch.exp(logp - old_logp)  # ppo.py:33
    surrogate = torch.min(ratio * adv, ratio.clamp(1 - clip_ratio, 1 + clip_ratio) * adv)  # ppo.py:34-36
    policy_loss = -surrogate.mean()
    entropy_loss = -entropy_coef * ent.mean()  # ppo.py:39
    vf_loss = vf_coef * (v - ret).pow(2).mean()  # ppo.py:42
    loss = policy_loss + entropy_loss + vf_loss  # ppo.py:45
    approx_kl = (old_logp - logp).mean()  # ppo.py:48
    clipped = ratio.lt(1 - clip_ratio) | ratio.gt(1 + clip_ratio)  # ppo.py:51-52
    info = dict(policy_loss=policy_loss.item(), entropy_loss=entropy_loss.item(), vf_loss=vf_loss.item(),
                total_loss=loss.item(), approx_kl=approx_kl.item(),
                clip_fraction=clipped.to(torch.float32).mean().item())
    return loss, info
