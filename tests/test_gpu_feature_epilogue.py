"""GPU parity of the hand-written glue of the bf16 feature stage (csrc/feature_epilogue.cu) against torch:
bit-exact (bf16 rounding of the same fp32 values)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,H,W,C", [(1, 1, 1, 64), (3, 5, 7, 64), (64, 64, 64, 64), (2, 16, 16, 128)])
def test_conv_bias_relu_in_place(B, H, W, C):
    from ocrl_b200 import abi

    g = torch.Generator().manual_seed(B * 131 + C)
    y = torch.randn(B, C, H, W, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(C, generator=g).cuda()
    ref = torch.relu(y.float() + bias.view(1, C, 1, 1)).to(torch.bfloat16)
    abi.check(abi.lib().ocrl_conv_bias_relu_bf16(ctypes.c_void_p(y.data_ptr()), abi.ptr(bias), B * H * W, C,
                                                 abi.stream_ptr()), "ocrl_conv_bias_relu_bf16")
    torch.cuda.synchronize()
    assert torch.equal(y, ref)


@pytest.mark.parametrize("B,C,S", [(1, 3, 4), (5, 3, 64), (2, 1, 33), (64, 3, 64)])
def test_frames_to_padded_nhwc(B, C, S):
    from ocrl_b200 import abi

    g = torch.Generator().manual_seed(B + 17 * C + S)
    obs = torch.rand(B, C, S, S, generator=g).cuda()
    out = torch.empty(B, 8, S, S, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    abi.check(abi.lib().ocrl_frames_to_nhwc_bf16(abi.ptr(obs), ctypes.c_void_p(out.data_ptr()), B, C, S, S, 8,
                                                 abi.stream_ptr()), "ocrl_frames_to_nhwc_bf16")
    torch.cuda.synchronize()
    assert torch.equal(out[:, :C], obs.to(torch.bfloat16))
    assert (out[:, C:] == 0).all()


def test_bad_shapes_are_reported():
    from ocrl_b200 import abi

    y = torch.zeros(8, 48, device="cuda", dtype=torch.bfloat16)
    b = torch.zeros(48, device="cuda")
    rc = abi.lib().ocrl_conv_bias_relu_bf16(ctypes.c_void_p(y.data_ptr()), abi.ptr(b), 8, 48, abi.stream_ptr())
    assert rc == -1 and b"channels=48" in abi.lib().ocrl_last_error()


@pytest.mark.parametrize("B,S", [(1, 16), (3, 32), (64, 64), (2, 128)])
def test_fused_first_convolution(B, S):
    """relu(conv5x5(obs) + b) of the first encoder layer: same operand rounding as bf16 autocast (bf16 frames and
    weights, fp32 accumulate), so the result equals an fp32 convolution of the rounded operands up to one bf16
    rounding of the output."""
    from ocrl_b200 import abi

    g = torch.Generator().manual_seed(B * 7 + S)
    obs = torch.rand(B, 3, S, S, generator=g).cuda()
    w = (0.3 * torch.randn(64, 3, 5, 5, generator=g)).cuda()
    b = (0.1 * torch.randn(64, generator=g)).cuda()
    out = torch.empty(B, 64, S, S, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    abi.check(abi.lib().ocrl_conv_first_relu_bf16(abi.ptr(obs), abi.ptr(w), abi.ptr(b), ctypes.c_void_p(out.data_ptr()),
                                                  B, 3, S, S, 64, abi.stream_ptr()), "ocrl_conv_first_relu_bf16")
    torch.cuda.synchronize()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    ref = torch.relu(torch.nn.functional.conv2d(obs.bfloat16().float(), w.bfloat16().float(), b, padding=2))
    torch.backends.cudnn.allow_tf32 = prev
    err = (out.float() - ref).abs().max() / ref.abs().max()
    assert err < 6e-3, float(err)
    assert float((out.float() - ref).norm() / ref.norm()) < 3e-3
