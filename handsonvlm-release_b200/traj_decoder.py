"""Trajectory head, generation side: drop-in for the reference's ``CVAETrajDecoder.inference`` with the gather and the
decoder MLP fused into one CUDA launch.

Mirrors ``handsonvlm/model/language_model/traj_decoder.py:7-70`` (``TrajDecoder`` / ``CVAETrajDecoder``) over
``hoi_forecast/architecture/traj_decoder.py:8-91`` (``TrajCVAE``) and ``decoder_modules.py:5-61`` (``VAE``): same
constructor argument, same module tree and therefore the same state-dict keys
(``hand_traj_decoder.cvae.{enc_MLP.0,linear_means,linear_log_var,dec_MLP.0,dec_MLP.2}.{weight,bias}``), same
``inference(pred_hand_embeddings=...)`` keyword call.  The noise is drawn exactly as the reference draws it
(``z_scale * torch.randn([R, latent], device=...)`` then cast, traj_decoder.py:87), so a seeded run reproduces the
reference's sample.

Only the generation path is built here; the training-side ``forward`` (CVAE encoder + reparameterisation + losses) is not
part of the visual-token path and raises ``NotImplementedError``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _VAE(nn.Module):
    """Parameter container with the reference VAE's layout (decoder_modules.py:5-30)."""

    def __init__(self, in_dim, hidden_dim, latent_dim, condition_dim):
        super().__init__()
        self.in_dim, self.latent_dim, self.condition_dim = in_dim, latent_dim, condition_dim
        self.enc_MLP = nn.Sequential(nn.Linear(in_dim + condition_dim, hidden_dim), nn.ELU())
        self.linear_means = nn.Linear(hidden_dim, latent_dim)
        self.linear_log_var = nn.Linear(hidden_dim, latent_dim)
        self.dec_MLP = nn.Sequential(nn.Linear(latent_dim + condition_dim, hidden_dim), nn.ELU(),
                                     nn.Linear(hidden_dim, in_dim))


class _TrajCVAE(nn.Module):
    """``TrajCVAE(condition_contact=False)`` (hoi_forecast/architecture/traj_decoder.py:8-31)."""

    def __init__(self, in_dim, hidden_dim, latent_dim, token_dim, z_scale=2.0):
        super().__init__()
        if in_dim != 2:
            raise ValueError("the fused trajectory head decodes 2-d hand coordinates (in_dim == 2)")
        self.latent_dim, self.token_dim, self.z_scale = latent_dim, token_dim, z_scale
        self.cvae = _VAE(in_dim, hidden_dim, latent_dim, token_dim)

    def _dec(self):
        d = self.cvae.dec_MLP
        return d[0].weight, d[0].bias, d[2].weight, d[2].bias

    def _noise(self, rows, like):
        # traj_decoder.py:87 -- fp32 randn on the device, scaled, then cast to the embedding dtype
        return (self.z_scale * torch.randn([rows, self.latent_dim], device=like.device)).to(like.dtype)

    @torch.no_grad()
    def inference(self, hand_embedding, contact_point=None, z=None):
        """[R, token_dim] -> [R, 2] (traj_decoder.py:75-91).  ``z`` overrides the internally drawn noise."""
        R = hand_embedding.shape[0]
        assert hand_embedding.shape == torch.Size([R, self.token_dim]), hand_embedding.shape
        if z is None:
            z = self._noise(R, hand_embedding)
        out = ops.traj_decode(hand_embedding, z, *self._dec(), interleaved=False)
        return out.to(hand_embedding.dtype)

    @torch.no_grad()
    def inference_step(self, hidden_last, z=None):
        """Raw last hidden rows [B, 2*token_dim] -> [2B, 2]; the even/odd split of handsonvlm.py:613-616 is fused."""
        B = hidden_last.shape[0]
        assert hidden_last.shape == torch.Size([B, 2 * self.token_dim]), hidden_last.shape
        if z is None:
            z = self._noise(2 * B, hidden_last)
        out = ops.traj_decode(hidden_last, z, *self._dec(), interleaved=True)
        return out.to(hidden_last.dtype)


class CVAETrajDecoder(nn.Module):
    """Drop-in for ``CVAETrajDecoder(token_dim)`` (handsonvlm/model/language_model/traj_decoder.py:60-70);
    ``token_dim`` is the per-hand width D/2 (handsonvlm.py:65)."""

    def __init__(self, token_dim, hidden_dim=512, latent_dim=256):
        super().__init__()
        self.token_dim = token_dim
        self.hand_traj_decoder = _TrajCVAE(in_dim=2, hidden_dim=hidden_dim, latent_dim=latent_dim, token_dim=token_dim)

    def forward(self, **kwargs):
        raise NotImplementedError("training-side CVAE (encoder, reparameterisation, losses) is outside the visual-token "
                                  "path this library replaces; keep the reference module for training")

    def inference(self, **kwargs):
        """``pred_hand_embeddings`` [B,2,T_pred,token_dim] -> [B,2,T_pred,2] (traj_decoder.py:39-47)."""
        e = kwargs["pred_hand_embeddings"]
        B, T_pred = e.shape[0], e.shape[2]
        assert e.shape == torch.Size([B, 2, T_pred, self.token_dim]), e.shape
        out = self.hand_traj_decoder.inference(e.reshape(-1, self.token_dim), z=kwargs.get("z"))
        return out.reshape(B, 2, T_pred, 2)

    def inference_step(self, hidden_last, z=None):
        """The sampling loop's ``<hand_traj>`` branch (handsonvlm.py:609-622) in one launch:
        ``hidden_states[-1][:, -1, :]`` [B, D] -> ``pred_hand`` [B, 2, 2]."""
        B = hidden_last.shape[0]
        return self.hand_traj_decoder.inference_step(hidden_last, z=z).reshape(B, 2, 2)


class _TrajMLP(nn.Module):
    """``TrajMLP`` parameters (hoi_forecast/architecture/traj_decoder.py:94-104)."""

    def __init__(self, hidden_dim, token_dim):
        super().__init__()
        self.token_dim = token_dim
        self.mlp = nn.Sequential(nn.Linear(token_dim, hidden_dim), nn.ReLU(inplace=True),
                                 nn.Linear(hidden_dim, hidden_dim), nn.ReLU(inplace=True), nn.Linear(hidden_dim, 2))

    @torch.no_grad()
    def inference(self, hand_embedding, contact_point=None):
        R = hand_embedding.shape[0]
        assert hand_embedding.shape == torch.Size([R, self.token_dim]), hand_embedding.shape
        m = self.mlp
        h = ops.skinny_linear(hand_embedding, m[0].weight, m[0].bias, "relu")
        h = ops.skinny_linear(h, m[2].weight, m[2].bias, "relu")
        return ops.skinny_linear(h, m[4].weight, m[4].bias, "none")


class MLPTrajDecoder(nn.Module):
    """Drop-in for ``MLPTrajDecoder(token_dim)`` (handsonvlm/model/language_model/traj_decoder.py:50-57), generation side:
    same state-dict keys (``hand_traj_decoder.mlp.{0,2,4}``), same ``inference(pred_hand_embeddings=...)`` call."""

    def __init__(self, token_dim, hidden_dim=512):
        super().__init__()
        self.token_dim = token_dim
        self.hand_traj_decoder = _TrajMLP(hidden_dim, token_dim)

    def forward(self, **kwargs):
        raise NotImplementedError("training-side losses are outside the visual-token path this library replaces; keep the "
                                  "reference module for training")

    def inference(self, **kwargs):
        e = kwargs["pred_hand_embeddings"]
        B, T_pred = e.shape[0], e.shape[2]
        assert e.shape == torch.Size([B, 2, T_pred, self.token_dim]), e.shape
        return self.hand_traj_decoder.inference(e.reshape(-1, self.token_dim)).reshape(B, 2, T_pred, 2)

    def inference_step(self, hidden_last):
        """``hidden_states[-1][:, -1, :]`` [B, D] -> ``pred_hand`` [B, 2, 2] (handsonvlm.py:609-622)."""
        B = hidden_last.shape[0]
        e = ops.hand_gather_step(hidden_last)                     # [B,2,1,D/2]
        return self.inference(pred_hand_embeddings=e).reshape(B, 2, 2)
