"""Drop-in ``SLATE_Module`` (reference: ocrs/slate/slate_module.py:23-267).

Same constructor (``ocr_config``, ``env_config`` attribute objects -- Hydra DictConfig or
SimpleNamespace), same sub-module names (checkpoint compatible), same ``forward`` / ``get_loss``
/ ``get_samples`` behaviour.  ``forward`` gets defaults for ``with_attns`` / ``with_masks`` so the
bare module also works behind ``sb3s/ocr_extractor.py:45``.

Hot path (``_get_slots``): cuDNN convs -> [fused kernel: +position table, NCHW->tokens,
LayerNorm+MLP, LayerNorm, k/v projection] -> fused T-iteration kernel.
"""
import os
from itertools import chain

import numpy as np
import torch
from torch import nn

from .adjacent import (BosToken, BroadCastDecoder, LearnedPositionalEncoding, OneHotDictionary,
                       TransformerDecoder, cosine_anneal, dVAE, gumbel_softmax)
from .feature_stage import (FusedBf16Encoder, PositionalEmbedding, SlotAttnCNNEncoder, frames_to_obs,
                            is_u8_frames)
from .networks import linear
from .slot_attn import SlotAttentionEncoder


def img_to_slot(x):  # [B,D,H,W] -> [B,N,D]
    return x.permute(0, 2, 3, 1).reshape(x.shape[0], -1, x.shape[1])


def for_viz(x):
    return np.array(x.clamp(0, 1).permute(0, 2, 3, 1).detach().cpu().numpy() * 255.0, dtype=np.uint8)


def visualize(images):
    tiles = []
    for img in images:
        tiles += [img] if img.dim() == 4 else list(torch.unbind(img, dim=1))
    return torch.cat(tiles, dim=-1)


def calculate_ari(true_masks, pred_masks):
    """utils/tools.py:309-320: ARI between argmax segmentations, per image, on the CPU."""
    from sklearn.metrics import adjusted_rand_score

    t = torch.argmax(true_masks.flatten(2), dim=1).detach().cpu().numpy()
    p = torch.argmax(pred_masks.flatten(2), dim=1).detach().cpu().numpy()
    return [adjusted_rand_score(t[b], p[b]) for b in range(t.shape[0])]


class SLATE_Module(nn.Module):
    def __init__(self, ocr_config, env_config) -> None:
        super().__init__()
        self._obs_size = obs_size = env_config.obs_size
        self._obs_channels = obs_channels = env_config.obs_channels
        self._use_cnn_feat = use_cnn_feat = ocr_config.use_cnn_feat
        self._use_bcdec = ocr_config.use_bcdec
        self._vocab_size = vocab_size = ocr_config.dvae.vocab_size
        self._d_model = d_model = ocr_config.dvae.d_model
        cnn_hsize = ocr_config.cnn.hidden_size
        sa = ocr_config.slotattr
        self._num_slots = num_slots = sa.num_slots
        slot_size = sa.slot_size
        dropout = ocr_config.learning.dropout
        self._tau_start = ocr_config.tau_start
        self._tau_final = ocr_config.tau_final
        self._tau_steps = ocr_config.tau_steps
        self._tau = 1.0
        self._hard = ocr_config.hard

        self._dvae = dVAE(vocab_size, obs_channels)
        self._enc_size = enc_size = obs_size // 4
        self._enc = SlotAttnCNNEncoder(obs_size, obs_channels, cnn_hsize)
        self._enc_pos = PositionalEmbedding(obs_size, cnn_hsize)
        self._slotattn = SlotAttentionEncoder(sa.num_iterations, num_slots, cnn_hsize, slot_size,
                                              sa.mlp_hidden_size, sa.pos_channels, sa.num_slot_heads,
                                              kv_dtype=getattr(sa, "kv_dtype", None))  # optional key of this implementation
        if self._use_bcdec:
            self._dec = BroadCastDecoder(obs_size, obs_channels, cnn_hsize, slot_size)
        self._slotproj = linear(slot_size, d_model, bias=False)
        self._dict = OneHotDictionary(vocab_size, d_model)
        self._bos_token = BosToken(d_model)
        self._z_pos = LearnedPositionalEncoding(1 + enc_size**2, d_model, dropout)
        self._tfdec = TransformerDecoder(ocr_config.tfdec.num_dec_blocks, enc_size**2, d_model,
                                         ocr_config.tfdec.num_dec_heads, dropout)
        self._out = linear(d_model, vocab_size, bias=False)

        if use_cnn_feat:
            self.num_slots = obs_size**2
            self.rep_dim = cnn_hsize + obs_channels
        else:
            self.num_slots = num_slots
            self.rep_dim = slot_size

    # ---- optimiser parameter groups (slate.py:19-34) ---------------------------------------------
    def get_dvae_params(self):
        return self._dvae.parameters()

    def get_sa_params(self):
        groups = [self._enc.parameters(), self._enc_pos.parameters(), self._slotattn.parameters(),
                  self._slotproj.parameters()]
        if self._use_bcdec:
            groups.append(self._dec.parameters())
        return chain(*groups)

    def get_tfdec_params(self):
        return chain(self._dict.parameters(), self._bos_token.parameters(), self._z_pos.parameters(),
                     self._tfdec.parameters(), self._out.parameters())

    # ---- hot path ---------------------------------------------------------------------------------
    def _hot_needs_grad(self, obs):
        if not torch.is_grad_enabled():
            return False
        hot = chain(self._enc.parameters(), self._enc_pos.parameters(), self._slotattn.parameters())
        return obs.requires_grad or any(p.requires_grad for p in hot)

    def _conv_mode(self):
        """Precision of the (cuDNN) CNN encoder in the inference fast path: ``fp32`` follows torch's
        backend flags on NCHW tensors; ``tf32`` / ``bf16`` run channels-last on the tensor cores and hand the
        token-major feature map to the token-stage kernel without a transpose.  Default: bf16 k/v -> bf16 convs
        (measured end to end on B200: slots 1.7e-3, masks 2.7e-3 relative, inside the 2e-2 bf16-mode budget)."""
        mode = os.environ.get("OCRL_CONV_DTYPE")
        if mode is None:
            mode = "bf16" if self._slotattn.slot_attention.kv_dtype == "bf16" else "fp32"
        if mode not in ("fp32", "tf32", "bf16"):
            raise ValueError(f"OCRL_CONV_DTYPE must be fp32, tf32 or bf16, got {mode}")
        return mode

    def _encode_features(self, obs):
        mode = self._conv_mode()
        if mode == "fp32":
            return self._enc(obs)
        x = obs.contiguous(memory_format=torch.channels_last)
        if mode == "bf16":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return self._enc(x)
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = True
        try:
            return self._enc(x)
        finally:
            torch.backends.cudnn.allow_tf32 = prev

    def _pos_table_with_bias(self, fast):
        """Position table [C, S*S] plus the last convolution's bias, cached until one of their parameters changes
        (the table is input-independent; the reference recomputes it per batch element, utils.py:28-33)."""
        cm = self._enc_pos.channels_map
        last = self._enc._encoder[3]
        key = tuple((p.data_ptr(), p._version) for p in (cm.weight, cm.bias, last.bias))
        cached = self.__dict__.get("_pos_table_cache")
        if cached is None or cached[0] != key:
            cached = (key, (self._enc_pos.table() + fast.last_bias.unsqueeze(1)).contiguous())
            self.__dict__["_pos_table_cache"] = cached
        return cached[1]

    def _slot_attention(self, obs):
        u8 = is_u8_frames(obs)  # uint8 HWC frames (utils/datasets.py:17): only the all-hand-written bf16 path ingests them
        if u8 and (self._hot_needs_grad(obs) or self._conv_mode() != "bf16"
                   or os.environ.get("OCRL_CONV_FUSED", "1") == "0"):
            obs = frames_to_obs(obs)
        if not self._hot_needs_grad(obs):
            # inference: position add + (transpose) + token MLP + projections fused into one kernel
            with torch.no_grad():
                if self._conv_mode() == "bf16" and os.environ.get("OCRL_CONV_FUSED", "1") != "0":
                    # fused bias+ReLU convolutions; the last conv's bias rides on the position table
                    # conv_impl: 'ocrl' (default) = all four layers hand-written (mma.sync first layer + tcgen05 implicit
                    # GEMM), 'cudnn' = library convolutions for layers 2-4
                    impl = getattr(self, "conv_impl", "ocrl")
                    fast = self.__dict__.get("_fast_enc")
                    if fast is None or fast._convs != impl:
                        fast = self.__dict__["_fast_enc"] = FusedBf16Encoder(self._enc, convs=impl)
                    return self._slotattn(fast(obs), _pos_table=self._pos_table_with_bias(fast))
                return self._slotattn(self._encode_features(obs), _pos_table=self._enc_pos.table())
        fmap = self._enc(obs)
        emb = self._enc_pos(fmap).permute(0, 2, 3, 1).flatten(start_dim=1, end_dim=2)
        return self._slotattn(emb)

    def _get_z(self, obs):
        z, z_logits = self._dvae(obs, self._tau, self._hard)
        z_hard = gumbel_softmax(z_logits, self._tau, True, dim=1).detach()
        return z, z_hard

    def _get_slots(self, obs, with_attns=False, z_hard=None, with_ce=False):
        slots, attns = self._slot_attention(obs)
        res = [slots]
        if with_attns:
            res.append(attns)
        if with_ce:
            z_hard = z_hard.permute(0, 2, 3, 1).flatten(start_dim=1, end_dim=2)
            z_emb = self._dict(z_hard)
            z_emb = torch.cat([self._bos_token().expand(obs.shape[0], -1, -1), z_emb], dim=1)
            z_emb = self._z_pos(z_emb)
            pred = self._out(self._tfdec(z_emb[:, :-1], self._slotproj(slots)))
            ce = -(z_hard * torch.log_softmax(pred, dim=-1)).flatten(start_dim=1).sum(-1).mean()
            res.append(ce)
        return res[0] if len(res) == 1 else tuple(res)

    def _gen_imgs(self, slots):
        memory = self._slotproj(slots)
        z_gen = memory.new_zeros(0)
        tokens = self._bos_token().expand(memory.shape[0], 1, -1)
        for _ in range(self._enc_size**2):
            dec = self._tfdec(self._z_pos(tokens), memory)
            z_next = torch.nn.functional.one_hot(self._out(dec)[:, -1:].argmax(dim=-1), self._vocab_size)
            z_gen = torch.cat((z_gen, z_next), dim=1)
            tokens = torch.cat([tokens, self._dict(z_next)], dim=1)
        z_gen = z_gen.transpose(1, 2).float().reshape(memory.shape[0], -1, self._enc_size, self._enc_size)
        return self._dvae.decode(z_gen)

    def _attn_maps(self, attns, obs, num_slots):
        h, w = (obs.shape[1], obs.shape[2]) if is_u8_frames(obs) else (obs.shape[2], obs.shape[3])
        return attns.transpose(-1, -2).reshape(obs.shape[0], num_slots, 1, h, w)

    def forward(self, obs, with_attns=False, with_masks=False):
        assert not (with_attns and with_masks)  # one of attns and masks can be returned
        if self._use_cnn_feat:
            return img_to_slot(torch.cat([self._enc_pos(self._enc(obs)), obs], dim=1))
        if is_u8_frames(obs) and (self._use_cnn_feat or with_attns):
            obs = frames_to_obs(obs)  # these outputs contain the float frames themselves
        if with_attns or with_masks:
            slots, attns = self._get_slots(obs, with_attns=True)
            attns = self._attn_maps(attns, obs, slots.shape[1])
            if with_attns:
                attns = obs.unsqueeze(1) * attns + (1.0 - attns)
            return slots, attns
        return self._get_slots(obs)

    def get_loss(self, obs, masks, with_rep=False, with_mse=False) -> dict:
        z, z_hard = self._get_z(obs)
        recon = self._dvae.decode(z)
        dvae_mse = ((obs - recon) ** 2).sum() / obs.shape[0]
        slots, attns, cross_entropy = self._get_slots(obs, z_hard=z_hard, with_attns=True, with_ce=True)
        attns = self._attn_maps(attns, obs, slots.shape[1])
        if masks is not None:
            fg_mask = 1 - masks[:, -1].unsqueeze(1)
            attns = torch.cat([attns * fg_mask, fg_mask], dim=1)
            ari = np.mean(calculate_ari(masks, attns))
        else:
            ari = 0
        if self._use_bcdec:
            recon = self._dec(slots)
            mse = ((obs - recon) ** 2).sum() / obs.shape[0]
            metrics = {"loss": mse, "mse": mse.detach(), "ari": ari}
        else:
            metrics = {"loss": dvae_mse + cross_entropy, "dvae_mse": dvae_mse.detach(),
                       "cross_entropy": cross_entropy.detach(), "tau": torch.Tensor([self._tau])}
            if with_mse:
                recon_tf = self._gen_imgs(slots)
                metrics["mse"] = (((obs - recon_tf) ** 2).sum() / obs.shape[0]).detach()
        return (metrics, z) if with_rep else metrics

    def get_samples(self, obs) -> dict:
        z, _ = self._get_z(obs)
        recon = self._dvae.decode(z)
        slots, attns = self._get_slots(obs, with_attns=True)
        attns = self._attn_maps(attns, obs, self._num_slots)
        attns = obs.unsqueeze(1) * attns + (1.0 - attns)
        if self._use_bcdec:
            return {"samples": for_viz(visualize([obs, self._dec(slots), attns]))}
        return {"samples": for_viz(visualize([obs, recon, self._gen_imgs(slots), attns]))}

    def update_tau(self, step: int) -> None:
        self._tau = cosine_anneal(step, self._tau_start, self._tau_final, 0, self._tau_steps)
