"""On-device MLP policy for the "PPO-policy actions" variant of the headline config.

Architecture of the agent the reference trains and ships
(examples/stable_baselines3/DeepRMSA.ipynb: ``PPO(MlpPolicy, env, policy_kwargs=dict(net_arch=5*[128]))``;
examples/stable_baselines3/bkp/deeprmsa-ppo-trained/best_model.zip): observation 54 -> 5 x (Linear 128 + tanh)
shared trunk -> ``action_net`` (5 logits) and ``value_net`` (1).  ``from_sb3_zip`` reads ``policy.pth`` out of
such an archive (a plain ``state_dict``; Stable-Baselines3 itself is not needed).

Two forward paths: :meth:`forward` (plain torch float32, the numerics reference) and :meth:`act_native` -- the whole
``model.predict(obs, deterministic=True)`` as ONE fused sm_100a kernel of liborlg.so (``orlg_policy_act``: bf16
``tcgen05.mma`` with TMEM accumulators, weights resident in shared memory, tanh + argmax epilogue, int32 actions out).
"""
from __future__ import annotations

import io
import zipfile
from typing import Sequence

import torch


class MlpPolicy(torch.nn.Module):
    def __init__(self, obs_dim: int = 54, n_actions: int = 5, net_arch: Sequence[int] = (128,) * 5):
        super().__init__()
        layers, d = [], obs_dim
        for width in net_arch:
            layers += [torch.nn.Linear(d, width), torch.nn.Tanh()]
            d = width
        self.shared_net = torch.nn.Sequential(*layers)
        self.action_net = torch.nn.Linear(d, n_actions)
        self.value_net = torch.nn.Linear(d, 1)
        self.critic_net = None          # the critic's own trunk when the checkpoint has separate ones (SB3 >= 1.8)

    @classmethod
    def from_state_dict(cls, sd) -> "MlpPolicy":
        """``policy.pth`` of an SB3 ``ActorCriticPolicy`` with a shared trunk (keys ``mlp_extractor.shared_net.<2i>``)
        or with separate actor/critic trunks (``mlp_extractor.policy_net`` / ``mlp_extractor.value_net``, SB3 >= 1.8: the
        actor's trunk becomes ``shared_net``, the critic's is kept as ``critic_net`` so that ``forward`` returns the right value;
        the fused kernel then evaluates the actor only and its value output is not meaningful)."""
        shared = any(k.startswith("mlp_extractor.shared_net.") for k in sd)
        trunk = "mlp_extractor.shared_net." if shared else "mlp_extractor.policy_net."
        idx = sorted({int(k[len(trunk):].split(".")[0]) for k in sd if k.startswith(trunk)})
        widths = [sd["%s%d.weight" % (trunk, i)].shape[0] for i in idx]
        obs_dim = sd["%s%d.weight" % (trunk, idx[0])].shape[1]
        pol = cls(obs_dim, sd["action_net.weight"].shape[0], widths)
        mine = {}
        for j, i in enumerate(idx):
            mine["shared_net.%d.weight" % (2 * j)] = sd["%s%d.weight" % (trunk, i)]
            mine["shared_net.%d.bias" % (2 * j)] = sd["%s%d.bias" % (trunk, i)]
        for k in ("action_net.weight", "action_net.bias", "value_net.weight", "value_net.bias"):
            mine[k] = sd[k]
        pol.load_state_dict(mine)
        vtrunk = "mlp_extractor.value_net."
        if not shared and any(k.startswith(vtrunk) for k in sd):
            vidx = sorted({int(k[len(vtrunk):].split(".")[0]) for k in sd if k.startswith(vtrunk)})
            layers = []
            for i in vidx:
                w = sd["%s%d.weight" % (vtrunk, i)]
                lin = torch.nn.Linear(w.shape[1], w.shape[0])
                lin.load_state_dict({"weight": w, "bias": sd["%s%d.bias" % (vtrunk, i)]})
                layers += [lin, torch.nn.Tanh()]
            pol.critic_net = torch.nn.Sequential(*layers)
        return pol

    @classmethod
    def from_sb3_zip(cls, path, device=None) -> "MlpPolicy":
        with zipfile.ZipFile(path) as z:
            sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
        pol = cls.from_state_dict(sd)
        return pol.to(device) if device is not None else pol

    # ------------------------------------------------------------------ fused native kernel (csrc/orlg_policy.cuh)
    def _native_handle(self, device_index: int):
        import ctypes as C

        import numpy as np

        from . import _native as nat

        key = ("h", device_index)
        cache = self.__dict__.setdefault("_native_cache", {})
        if key not in cache:
            lins = [m for m in self.shared_net if isinstance(m, torch.nn.Linear)]
            mats = lins + [self.action_net, self.value_net]
            w = np.concatenate([m.weight.detach().float().cpu().numpy().reshape(-1) for m in mats]).astype(np.float32)
            b = np.concatenate([m.bias.detach().float().cpu().numpy().reshape(-1) for m in mats]).astype(np.float32)
            h = C.c_void_p()
            nat.check(nat.lib().orlg_policy_create(device_index, lins[0].in_features, lins[0].out_features, len(lins),
                                                   self.action_net.out_features, w.ctypes.data_as(C.c_void_p),
                                                   b.ctypes.data_as(C.c_void_p), C.byref(h)))
            cache[key] = h
        return cache[key]

    @torch.no_grad()
    def act_native(self, obs: torch.Tensor, out: torch.Tensor = None, logits: torch.Tensor = None) -> torch.Tensor:
        """Deterministic actions int32 ``[N, 1]`` from float32 CUDA observations ``[N, obs_dim]`` through the fused tensor-core
        kernel; ``logits`` (optional float32 ``[N, n_actions + 1]``) receives the logits and the value estimate."""
        import ctypes as C

        from . import _native as nat

        assert obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous()
        n = obs.shape[0]
        if out is None:
            out = torch.empty((n, 1), dtype=torch.int32, device=obs.device)
        h = self._native_handle(obs.device.index)
        stream = C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream)
        nat.check(nat.lib().orlg_policy_act(h, C.c_void_p(obs.data_ptr()), n, C.c_void_p(out.data_ptr()),
                                            None if logits is None else C.c_void_p(logits.data_ptr()), stream))
        return out

    def forward(self, obs: torch.Tensor):
        """(logits [N, n_actions], value [N]) -- observations are used as they are (SB3 ``preprocess_obs`` of a
        Box space is a cast to float)."""
        x = obs.to(self.action_net.weight.dtype)
        latent = self.shared_net(x)
        latent_vf = latent if self.critic_net is None else self.critic_net(x)
        return self.action_net(latent), self.value_net(latent_vf).squeeze(-1)

    @torch.no_grad()
    def act(self, obs: torch.Tensor, deterministic: bool = True, generator=None) -> torch.Tensor:
        """int32 actions [N, 1] ready for ``OpticalVecEnv.step_raw`` (``model.predict(obs, deterministic=...)``)."""
        logits, _ = self.forward(obs)
        if deterministic:
            a = logits.argmax(dim=-1)
        else:
            a = torch.multinomial(torch.softmax(logits.float(), dim=-1), 1, generator=generator).squeeze(-1)
        return a.to(torch.int32).unsqueeze(-1)
