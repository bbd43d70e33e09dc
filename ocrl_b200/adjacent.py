"""Training-step neighbours of the hot path, kept in plain torch (library GEMMs / cuDNN):
the discrete VAE, the SLATE transformer decoder and the spatial-broadcast decoder.

They are NOT on the slot-attention hot path but
``SLATE_Module.get_loss`` / ``get_samples`` need them, and their parameter names are part of the
checkpoint contract.  Behaviour follows ocrs/common/models.py:10-45,110-141,
ocrs/common/transformer.py:7-226 and ocrs/common/utils.py:75-85.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .feature_stage import PositionalEmbedding
from .networks import Conv2dBlock, conv2d, linear


def gumbel_softmax(logits, tau=1.0, hard=False, dim=-1):
    tiny = torch.finfo(logits.dtype).tiny
    g = -(torch.empty_like(logits).exponential_() + tiny).log()
    y = F.softmax((logits + g) / tau, dim)
    if not hard:
        return y
    idx = y.argmax(dim, keepdim=True)
    one_hot = torch.zeros_like(logits).scatter_(dim, idx, 1.0)
    return one_hot - y.detach() + y


def cosine_anneal(step, start_value, final_value, start_step, final_step):
    import math

    assert start_value >= final_value and start_step <= final_step
    if step < start_step:
        return start_value
    if step >= final_step:
        return final_value
    half_span, mid = 0.5 * (start_value - final_value), 0.5 * (start_value + final_value)
    return half_span * math.cos(math.pi * (step - start_step) / (final_step - start_step)) + mid


def linear_warmup(step, start_value, final_value, start_step, final_step):
    assert start_value <= final_value and start_step <= final_step
    if step < start_step:
        return start_value
    if step >= final_step:
        return final_value
    return (final_value - start_value) * (step + 1 - start_step) / (final_step - start_step) + start_value


class dVAE(nn.Module):
    def __init__(self, vocab_size, img_channels):
        super().__init__()
        enc = [Conv2dBlock(img_channels, 64, 4, 4)] + [Conv2dBlock(64, 64, 1, 1) for _ in range(6)]
        self._encoder = nn.Sequential(*enc, conv2d(64, vocab_size, 1))

        def up():
            return [Conv2dBlock(64, 64, 3, 1, 1), Conv2dBlock(64, 64, 1, 1), Conv2dBlock(64, 64, 1, 1),
                    Conv2dBlock(64, 64 * 2 * 2, 1), nn.PixelShuffle(2)]

        self._decoder = nn.Sequential(Conv2dBlock(vocab_size, 64, 1), *up(), *up(), conv2d(64, img_channels, 1))

    def forward(self, obs, tau=1.0, hard=True):
        z_logits = F.log_softmax(self._encoder(obs), dim=1)
        return gumbel_softmax(z_logits, tau, hard, dim=1), z_logits

    def decode(self, z):
        return self._decoder(z)


class BroadCastDecoder(nn.Module):
    def __init__(self, obs_size, obs_channels, hidden_size, slot_size):
        super().__init__()
        self._obs_size = obs_size
        self._obs_channels = obs_channels
        self._decoder = nn.Sequential(
            Conv2dBlock(slot_size, hidden_size, 5, 1, 2),
            Conv2dBlock(hidden_size, hidden_size, 5, 1, 2),
            Conv2dBlock(hidden_size, hidden_size, 5, 1, 2),
            conv2d(hidden_size, obs_channels + 1, 3, 1, 1),
        )
        self._pos_emb = PositionalEmbedding(obs_size, slot_size)

    def forward(self, slots):
        B, K, D = slots.shape
        S = self._obs_size
        grid = slots.reshape(B * K, D, 1, 1).expand(B * K, D, S, S)
        out = self._decoder(self._pos_emb(grid))
        rgb = out[:, : self._obs_channels].view(B, K, self._obs_channels, S, S)
        alpha = out[:, -1:].view(B, K, 1, S, S).softmax(dim=1)
        return (rgb * alpha).sum(dim=1)


class MultiHeadAttention(nn.Module):
    def __init__(self, d_model, num_heads, dropout=0.0, gain=1.0):
        super().__init__()
        assert d_model % num_heads == 0, "d_model must be divisible by num_heads"
        self.d_model = d_model
        self.num_heads = num_heads
        self.attn_dropout = nn.Dropout(dropout)
        self.output_dropout = nn.Dropout(dropout)
        self.proj_q = linear(d_model, d_model, bias=False)
        self.proj_k = linear(d_model, d_model, bias=False)
        self.proj_v = linear(d_model, d_model, bias=False)
        self.proj_o = linear(d_model, d_model, bias=False, gain=gain)

    def forward(self, q, k, v, attn_mask=None):
        B, T, _ = q.shape
        S = k.shape[1]
        h = self.num_heads
        q = self.proj_q(q).view(B, T, h, -1).transpose(1, 2)
        k = self.proj_k(k).view(B, S, h, -1).transpose(1, 2)
        v = self.proj_v(v).view(B, S, h, -1).transpose(1, 2)
        scores = (q * q.shape[-1] ** -0.5) @ k.transpose(-1, -2)
        if attn_mask is not None:
            scores = scores.masked_fill(attn_mask, float("-inf"))
        probs = self.attn_dropout(F.softmax(scores, dim=-1))
        out = (probs @ v).transpose(1, 2).reshape(B, T, -1)
        return self.output_dropout(self.proj_o(out))


class LearnedPositionalEncoding(nn.Module):
    def __init__(self, max_len, d_model, dropout=0.1):
        super().__init__()
        self.dropout = nn.Dropout(dropout)
        self.pe = nn.Parameter(torch.zeros(1, max_len, d_model), requires_grad=True)
        nn.init.trunc_normal_(self.pe)

    def forward(self, x):
        return self.dropout(x + self.pe[:, : x.shape[1]])


class TransformerDecoderBlock(nn.Module):
    def __init__(self, max_len, d_model, num_heads, dropout=0.0, gain=1.0, is_first=False):
        super().__init__()
        self.is_first = is_first
        self.self_attn_layer_norm = nn.LayerNorm(d_model)
        self.self_attn = MultiHeadAttention(d_model, num_heads, dropout, gain)
        causal = torch.triu(torch.ones((max_len, max_len), dtype=torch.bool), diagonal=1)
        self.self_attn_mask = nn.Parameter(causal, requires_grad=False)
        self.encoder_decoder_attn_layer_norm = nn.LayerNorm(d_model)
        self.encoder_decoder_attn = MultiHeadAttention(d_model, num_heads, dropout, gain)
        self.ffn_layer_norm = nn.LayerNorm(d_model)
        self.ffn = nn.Sequential(
            linear(d_model, 4 * d_model, weight_init="kaiming"),
            nn.ReLU(),
            linear(4 * d_model, d_model, gain=gain),
            nn.Dropout(dropout),
        )

    def forward(self, x, memory):
        T = x.shape[1]
        mask = self.self_attn_mask[:T, :T]
        if self.is_first:  # the first block normalises the residual stream itself
            x = self.self_attn_layer_norm(x)
            x = x + self.self_attn(x, x, x, mask)
        else:
            y = self.self_attn_layer_norm(x)
            x = x + self.self_attn(y, y, y, mask)
        y = self.encoder_decoder_attn_layer_norm(x)
        x = x + self.encoder_decoder_attn(y, memory, memory)
        return x + self.ffn(self.ffn_layer_norm(x))


class TransformerDecoder(nn.Module):
    def __init__(self, num_blocks, max_len, d_model, num_heads, dropout=0.0):
        super().__init__()
        if num_blocks > 0:
            gain = (3 * num_blocks) ** (-0.5)
            self.blocks = nn.ModuleList(
                [TransformerDecoderBlock(max_len, d_model, num_heads, dropout, gain, is_first=(i == 0))
                 for i in range(num_blocks)])
        else:
            self.blocks = nn.ModuleList()
        self.layer_norm = nn.LayerNorm(d_model)

    def forward(self, x, memory):
        for block in self.blocks:
            x = block(x, memory)
        return self.layer_norm(x)


class OneHotDictionary(nn.Module):
    def __init__(self, vocab_size, emb_size):
        super().__init__()
        self.dictionary = nn.Embedding(vocab_size, emb_size)

    def forward(self, x):  # x: [B, N, vocab] one-hot (or soft) -> embeddings of the argmax token
        return self.dictionary(torch.argmax(x, dim=-1))


class BosToken(nn.Module):
    def __init__(self, d_model):
        super().__init__()
        self._bos_token = nn.Parameter(torch.Tensor(1, 1, d_model))
        nn.init.xavier_uniform_(self._bos_token)

    def forward(self):
        return self._bos_token
