"""Per-image feature stage in front of slot attention (reference: ocrs/common/models.py:96-107,
ocrs/common/utils.py:10-33): 4x conv5x5 (ReLU on the first three) and the linear-ramp position
embedding.  Same module/parameter names as the reference.

In bf16 mode the four convolutions are hand-written kernels (csrc/conv_first.cu: mma.sync first layer;
csrc/conv_tc.cu: tcgen05 implicit GEMM for the 64 -> 64 layers) that hand a padded channels-last feature map to the
token-stage kernel; the fp32 parity mode and the training path call cuDNN through torch.  The position-table add and
the NCHW -> token-major transpose are fused into the token-stage kernel.
"""
import torch
import torch.nn as nn

from .networks import Conv2dBlock, conv2d


class SlotAttnCNNEncoder(nn.Module):
    def __init__(self, obs_size, obs_channels, hidden_size):
        super().__init__()
        self._encoder = nn.Sequential(
            Conv2dBlock(obs_channels, hidden_size, 5, 1, 2),
            Conv2dBlock(hidden_size, hidden_size, 5, 1, 2),
            Conv2dBlock(hidden_size, hidden_size, 5, 1, 2),
            conv2d(hidden_size, hidden_size, 5, 1, 2),
        )

    def forward(self, obs):
        return self._encoder(obs)


def is_u8_frames(obs) -> bool:
    """uint8 HWC frames [B,H,W,3] as the datasets / environments hold them (utils/datasets.py:17)."""
    return obs.dtype == torch.uint8 and obs.dim() == 4 and obs.shape[-1] == 3


def frames_to_obs(frames):
    """The reference's ingest (utils/datasets.py:17): uint8 HWC -> float CHW in [0, 1]."""
    return frames.permute(0, 3, 1, 2).float() / 255.0


class PaddedMap:
    """A bf16 feature map in the padded channels-last layout of ocrl_conv5x5_c64_tc (include/ocrl_sa.h):
    ``data`` is the flat position array [(2 + B (H + 2)) (W + 4), 64]."""

    def __init__(self, data, B, H, W):
        self.data, self.B, self.H, self.W = data, B, H, W
        self.shape = (B, data.shape[1], H, W)
        self.is_cuda, self.device, self.dtype = data.is_cuda, data.device, data.dtype
        self.requires_grad = False

    def dim(self):
        return 4

    def to_nchw(self):
        """[B, 64, H, W] view-copy (tests / debugging)."""
        rows = self.data.view(-1, self.W + 4, self.data.shape[1])[2:]
        rows = rows.view(self.B, self.H + 2, self.W + 4, -1)[:, : self.H, 2: self.W + 2]
        return rows.permute(0, 3, 1, 2)


class FusedBf16Encoder:
    """Inference fast path of SlotAttnCNNEncoder in bf16: channels-last tensors, weights cast once and cached until
    a parameter changes, the 3 input channels zero-padded to 8 so that the first layer also takes a tensor-core
    kernel.  The convolutions are cuDNN (library calls) with bias and ReLU fused (``cudnn_convolution_relu``); the
    frame ingest (fp32 NCHW -> padded bf16 NHWC) is one hand-written kernel.  The last layer's bias is not applied
    here: ``last_bias`` is folded into the position table by the caller (the token-stage kernel adds that table
    anyway).  ``OCRL_CONV_EPILOGUE=ocrl`` runs plain convolutions followed by the hand-written in-place bias+ReLU
    pass instead (measured on B200: 82 + 11 us per layer against 94 us fused -- cuDNN picks a slower tile shape
    for the plain convolution, so the fused call wins by a hair end to end: 0.584 vs 0.612 ms per step)."""

    def __init__(self, enc: "SlotAttnCNNEncoder", convs: str = "ocrl"):
        """convs: 'ocrl' = hand-written kernels for all four layers where the shape allows (3 -> 64 -> 64 channels,
        frame width 32 / 64 / 128), returning a PaddedMap; 'cudnn' = the library path for layers 2-4."""
        self._enc = enc
        self._key = None
        self._w = None
        self._b = None
        self._convs = convs
        self._packed = None

    def _refresh(self):
        convs = [self._enc._encoder[i].m for i in range(3)] + [self._enc._encoder[3]]
        key = tuple((c.weight.data_ptr(), c.weight._version, c.bias.data_ptr(), c.bias._version) for c in convs)
        if key == self._key:
            return
        ws, bs = [], []
        for i, c in enumerate(convs):
            w = c.weight.detach()
            if i == 0 and w.shape[1] % 8:
                w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 8 - w.shape[1] % 8))
            ws.append(w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
            bs.append(c.bias.detach().float().contiguous())
        self._w, self._b, self._key = ws, bs, key
        self._b16 = [b.to(torch.bfloat16) for b in bs]
        self._w0_f32 = convs[0].weight.detach().float().contiguous()  # the hand-written first layer rounds it itself
        self._packed = None  # bf16 [25][64][64] copies for the tcgen05 layers, rebuilt on demand

    def _packed_weights(self):
        import ctypes

        from . import abi

        if self._packed is None:
            convs = [self._enc._encoder[i].m for i in range(1, 3)] + [self._enc._encoder[3]]
            out = []
            for c in convs:
                w = c.weight.detach().float().contiguous()
                pk = torch.empty(25 * 64 * 64, device=w.device, dtype=torch.bfloat16)
                abi.check(abi.lib().ocrl_conv5x5_pack_weights(abi.ptr(w), ctypes.c_void_p(pk.data_ptr()), w.shape[0], w.shape[1],
                                                              abi.stream_ptr()), "ocrl_conv5x5_pack_weights")
                out.append(pk)
            self._packed = out
        return self._packed

    def _own_path_ok(self, obs):
        if is_u8_frames(obs):
            B, H, W, C = obs.shape
        else:
            B, C, H, W = obs.shape
        hidden = [self._enc._encoder[i].m for i in range(3)] + [self._enc._encoder[3]]
        return (self._convs == "ocrl" and C == 3 and W in (32, 64, 128) and (H * W) % 128 == 0
                and (obs.dtype == torch.float32 or is_u8_frames(obs)) and obs.is_contiguous()
                and all(tuple(c.weight.shape[2:]) == (5, 5) and c.weight.shape[0] == 64 for c in hidden)
                and all(c.weight.shape[1] == 64 for c in hidden[1:]))

    def _own_path(self, obs):
        """All four layers hand-written: conv_first (mma.sync) -> three tcgen05 implicit-GEMM layers, activations in the
        padded channels-last layout (fresh buffers per call: nothing to keep consistent between calls or CUDA graphs)."""
        import ctypes

        from . import abi

        L = abi.lib()
        u8 = is_u8_frames(obs)
        if u8:
            B, H, W, C = obs.shape
        else:
            B, C, H, W = obs.shape
        nbytes = L.ocrl_conv_padded_bytes(B, H, W)
        bufs = [torch.empty(nbytes // 128, 64, device=obs.device, dtype=torch.bfloat16) for _ in range(2)]
        pk = self._packed_weights()
        st = abi.stream_ptr()
        first = L.ocrl_conv_first_relu_u8p if u8 else L.ocrl_conv_first_relu_bf16p  # uint8 HWC frames: `/ 255` on the way in
        abi.check(first(abi.ptr(obs), abi.ptr(self._w0_f32), abi.ptr(self._b[0]),
                        ctypes.c_void_p(bufs[0].data_ptr()), B, C, H, W, 64, st), "ocrl_conv_first_relu")
        src = 0
        for i in range(3):  # layers 2, 3 (bias + ReLU) and 4 (bias rides on the position table, no ReLU)
            last = i == 2
            abi.check(L.ocrl_conv5x5_c64_tc(ctypes.c_void_p(bufs[src].data_ptr()), ctypes.c_void_p(pk[i].data_ptr()),
                                            None if last else abi.ptr(self._b[i + 1]),
                                            ctypes.c_void_p(bufs[1 - src].data_ptr()), B, H, W, 0 if last else 1, st),
                      "ocrl_conv5x5_c64_tc")
            src = 1 - src
        return PaddedMap(bufs[src], B, H, W)

    @property
    def last_bias(self):
        self._refresh()
        return self._b[3]

    def __call__(self, obs):
        import ctypes
        import os

        from . import abi

        self._refresh()
        if not obs.is_contiguous():  # e.g. a permuted view of HWC frames: the kernels read dense NCHW / HWC frames
            obs = obs.contiguous()
        if self._own_path_ok(obs):
            return self._own_path(obs)
        if is_u8_frames(obs):
            obs = frames_to_obs(obs)
        B, C, H, W = obs.shape
        cin = self._w[0].shape[1]
        use_cudnn_epilogue = os.environ.get("OCRL_CONV_EPILOGUE", "cudnn") == "cudnn"
        L = abi.lib()
        first = 0
        own_first = (os.environ.get("OCRL_CONV_FIRST", "ocrl") == "ocrl" and C == 3 and W % 16 == 0 and W <= 512
                     and self._w0_f32.shape[0] == 64 and tuple(self._w0_f32.shape[2:]) == (5, 5)
                     and obs.dtype == torch.float32 and obs.is_contiguous())
        if own_first:
            # hand-written fused first layer: fp32 NCHW frames in, relu(conv + bias) out as bf16 channels-last
            x = torch.empty((B, 64, H, W), device=obs.device, dtype=torch.bfloat16, memory_format=torch.channels_last)
            abi.check(L.ocrl_conv_first_relu_bf16(abi.ptr(obs), abi.ptr(self._w0_f32), abi.ptr(self._b[0]),
                                                  ctypes.c_void_p(x.data_ptr()), B, C, H, W, 64, abi.stream_ptr()),
                      "ocrl_conv_first_relu_bf16")
            first = 1
        elif cin == 8 and obs.dtype == torch.float32 and obs.is_contiguous():
            x = torch.empty((B, cin, H, W), device=obs.device, dtype=torch.bfloat16, memory_format=torch.channels_last)
            abi.check(L.ocrl_frames_to_nhwc_bf16(abi.ptr(obs), ctypes.c_void_p(x.data_ptr()), B, C, H, W, cin,
                                                 abi.stream_ptr()), "ocrl_frames_to_nhwc_bf16")
        else:
            x = torch.zeros((B, cin, H, W), device=obs.device, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
            x[:, :C] = obs
        # OCRL_CUDNN_BENCHMARK=1 lets cuDNN time its candidates once per shape (no gain measured on B200)
        autotune = os.environ.get("OCRL_CUDNN_BENCHMARK", "0") != "0"
        with torch.backends.cudnn.flags(enabled=True, benchmark=autotune, deterministic=False, allow_tf32=True):
            for i in range(first, 3):
                if use_cudnn_epilogue or self._w[i].shape[0] not in (64, 128):
                    x = torch.cudnn_convolution_relu(x, self._w[i], self._b16[i], [1, 1], [2, 2], [1, 1], 1)
                else:
                    x = torch.conv2d(x, self._w[i], None, 1, 2)
                    assert x.is_contiguous(memory_format=torch.channels_last)
                    abi.check(L.ocrl_conv_bias_relu_bf16(ctypes.c_void_p(x.data_ptr()), abi.ptr(self._b[i]),
                                                         B * x.shape[2] * x.shape[3], x.shape[1], abi.stream_ptr()),
                              "ocrl_conv_bias_relu_bf16")
            return torch.conv2d(x, self._w[3], None, 1, 2)


class PositionalEmbedding(nn.Module):
    def __init__(self, obs_size: int, obs_channels: int):
        super().__init__()
        ramp = torch.linspace(0, 1, obs_size)
        rev = torch.linspace(1, 0, obs_size)
        east = ramp.view(1, -1).expand(obs_size, obs_size)
        west = rev.view(1, -1).expand(obs_size, obs_size)
        south = ramp.view(-1, 1).expand(obs_size, obs_size)
        north = rev.view(-1, 1).expand(obs_size, obs_size)
        grid = torch.stack([north, south, west, east], dim=0).unsqueeze(0).contiguous()
        self.channels_map = nn.Conv2d(4, obs_channels, kernel_size=1)
        self.register_buffer("linear_position_embedding", grid)

    def table(self):
        """The input-independent [C, S*S] table the reference recomputes per batch element."""
        w = self.channels_map.weight.flatten(1)
        g = self.linear_position_embedding[0].flatten(1)
        return w @ g + self.channels_map.bias.unsqueeze(1)

    def forward(self, x):
        B, _, H, W = x.shape
        return x + self.table().view(1, -1, H, W)
