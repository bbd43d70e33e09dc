"""Timing of the fused first convolution against cuDNN's (bf16, channels-last, padded to 8 input channels)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrl_b200 import abi

def timeit(fn, warm=5, rep=30):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rep): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / rep * 1e3

B, S = 64, 64
obs = torch.rand(B, 3, S, S, device="cuda")
w = 0.3 * torch.randn(64, 3, 5, 5, device="cuda"); b = 0.1 * torch.randn(64, device="cuda")
out = torch.empty(B, 64, S, S, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
L = abi.lib()
f = lambda: L.ocrl_conv_first_relu_bf16(abi.ptr(obs), abi.ptr(w), abi.ptr(b), ctypes.c_void_p(out.data_ptr()), B, 3, S, S, 64, abi.stream_ptr())
print(f"ocrl fused first conv: {timeit(f):.1f} us  (output {out.numel()*2/1e6:.1f} MB)")
x8 = torch.zeros(B, 8, S, S, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last); x8[:, :3] = obs
w8 = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 5)).bfloat16().contiguous(memory_format=torch.channels_last)
b16 = b.bfloat16()
g = lambda: torch.cudnn_convolution_relu(x8, w8, b16, [1, 1], [2, 2], [1, 1], 1)
print(f"cuDNN conv+bias+relu (Cin padded to 8): {timeit(g):.1f} us")
x64 = torch.randn(B, 64, S, S, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
w64 = torch.randn(64, 64, 5, 5, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
h = lambda: torch.cudnn_convolution_relu(x64, w64, b16, [1, 1], [2, 2], [1, 1], 1)
print(f"cuDNN conv+bias+relu 64->64: {timeit(h):.1f} us")
h2 = lambda: torch.conv2d(x64, w64, None, 1, 2)
print(f"cuDNN conv 64->64 (no bias): {timeit(h2):.1f} us")
