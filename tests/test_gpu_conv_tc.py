"""GPU parity of the hand-written tcgen05 convolution (csrc/conv_tc.cu: layers 2-4 of SlotAttnCNNEncoder,
ocrs/common/models.py:96-107) and of the whole hand-written feature stage against the oracle's CNN encoder."""
import ctypes

import pytest
import torch

from ocrl_b200 import abi
from ocrl_b200.feature_stage import PaddedMap
from tests.golden_io import rel_err

pytestmark = pytest.mark.gpu


def _to_padded(x):  # [B,64,H,W] float -> padded channels-last bf16 array
    B, C, H, W = x.shape
    buf = torch.zeros(2 + B * (H + 2), W + 4, C, dtype=torch.bfloat16, device=x.device)
    view = buf[2:].view(B, H + 2, W + 4, C)
    view[:, :H, 2:W + 2] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    return buf.view(-1, C)


# (from ~300 tiles up the launcher takes the paired-tap kernel: two filter taps per N = 128 instruction, tiles of 127 positions)
@pytest.mark.parametrize("B,H,W", [(1, 8, 32), (3, 32, 32), (2, 64, 64), (5, 64, 64), (1, 128, 128), (2, 6, 64),
                                   (9, 64, 64), (20, 64, 64), (40, 32, 32), (33, 24, 32)])
@pytest.mark.parametrize("relu", [0, 1])
def test_conv5x5_matches_torch(B, H, W, relu):
    """bf16 operands, fp32 accumulation: compare with torch's conv2d on the same bf16-rounded operands in fp32."""
    torch.manual_seed(B * 100 + H + relu)
    dev = "cuda"
    x = torch.randn(B, 64, H, W, device=dev).to(torch.bfloat16).float()
    w = (torch.randn(64, 64, 5, 5, device=dev) * 0.05)
    bias = torch.randn(64, device=dev)
    L = abi.lib()
    pk = torch.empty(25 * 64 * 64, device=dev, dtype=torch.bfloat16)
    abi.check(L.ocrl_conv5x5_pack_weights(abi.ptr(w), ctypes.c_void_p(pk.data_ptr()), 64, 64, abi.stream_ptr()), "pack")
    xin = _to_padded(x)
    assert xin.numel() * 2 == L.ocrl_conv_padded_bytes(B, H, W)
    out = torch.full_like(xin, float("nan"))  # the kernel must write every position
    abi.check(L.ocrl_conv5x5_c64_tc(ctypes.c_void_p(xin.data_ptr()), ctypes.c_void_p(pk.data_ptr()), abi.ptr(bias),
                                    ctypes.c_void_p(out.data_ptr()), B, H, W, relu, abi.stream_ptr()), "conv")
    torch.cuda.synchronize()
    torch.backends.cudnn.allow_tf32 = False
    ref = torch.nn.functional.conv2d(x, w.to(torch.bfloat16).float(), bias, 1, 2)
    if relu:
        ref = ref.relu()
    got = PaddedMap(out, B, H, W).to_nchw().float()
    assert torch.isfinite(out.float()).all()
    err = rel_err(got.cpu(), ref.cpu())
    assert err < 6e-3, err  # bf16 rounding of the output
    # every padding position of the output is zero
    full = out.view(-1, W + 4, 64).float()
    assert float(full[:2].abs().max()) == 0 and float(full[:, :2].abs().max()) == 0 and float(full[:, W + 2:].abs().max()) == 0
    pad_rows = full[2:].view(B, H + 2, W + 4, 64)[:, H:]
    assert float(pad_rows.abs().max()) == 0


@pytest.mark.parametrize("size,B", [(64, 4), (32, 3)])
def test_hand_written_feature_stage_matches_oracle_cnn(size, B):
    """conv_first (mma.sync) + 3 x tcgen05 layers against the oracle's fp32 CNN encoder (last bias excluded: it rides
    on the position table); 2e-2 = the north-star bf16 tolerance."""
    import ocrl_b200
    from ocrl_b200.config import slate_config
    from ocrl_b200.feature_stage import FusedBf16Encoder
    from oracle import slot_oracle as so

    torch.manual_seed(3)
    model = ocrl_b200.SLATE(*slate_config(obs_size=size))
    model.to("cuda")
    enc = model._module._enc
    fast = FusedBf16Encoder(enc, convs="ocrl")
    obs = torch.rand(B, 3, size, size, device="cuda")
    fm = fast(obs)
    assert isinstance(fm, PaddedMap)
    p = {k: v.detach().cpu() for k, v in model._module.state_dict().items()}
    ref = so.cnn_encoder(obs.cpu(), p) - p["_enc._encoder.3.bias"].view(1, -1, 1, 1)
    err = rel_err(fm.to_nchw().float().cpu(), ref)
    print(f"hand-written feature stage vs fp32 oracle: {err:.2e}")
    assert err < 2e-2, err
    # and the library path of the same object agrees
    lib_map = FusedBf16Encoder(enc, convs="cudnn")(obs)
    assert rel_err(fm.to_nchw().float().cpu(), lib_map.float().cpu()) < 2e-2


@pytest.mark.parametrize("size,B", [(64, 5), (32, 2), (128, 1)])
def test_uint8_hwc_frames_are_ingested_bit_exactly(size, B):
    """uint8 HWC frames (utils/datasets.py:17) through ocrl_conv_first_relu_u8p: the `/ 255.0` of the reference's ingest
    happens inside the first convolution, bit-identical to feeding the converted float CHW tensor -- feature map, slots
    and masks (SLATE.__call__ takes the frames as they are in the bf16 mode)."""
    import ocrl_b200
    from ocrl_b200 import synth
    from ocrl_b200.config import slate_config
    from ocrl_b200.feature_stage import FusedBf16Encoder, frames_to_obs

    torch.manual_seed(5)
    model = ocrl_b200.SLATE(*slate_config(obs_size=size, kv_dtype="bf16"))
    model.to("cuda")
    model.eval()
    frames = torch.from_numpy(synth.random_objs_frames(B, size, seed=3)).cuda()  # [B,H,W,3] uint8
    assert frames.dtype == torch.uint8 and frames.shape == (B, size, size, 3)
    fast = FusedBf16Encoder(model._module._enc, convs="ocrl")
    a = fast(frames)
    b = fast(frames_to_obs(frames).contiguous())
    assert torch.equal(a.data, b.data)
    with torch.no_grad():
        torch.manual_seed(9)
        s8, m8 = model(frames, with_masks=True)
        torch.manual_seed(9)
        sf, mf = model(frames_to_obs(frames).contiguous(), with_masks=True)
    assert torch.equal(s8, sf) and torch.equal(m8, mf) and m8.shape == (B, 6, 1, size, size)
