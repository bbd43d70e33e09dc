"""Time the cuDNN CNN encoder under different precisions (what bounds end-to-end encode)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocrl_b200
from ocrl_b200.config import slate_config

def timeit(fn, warm=5, rep=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rep): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / rep

torch.backends.cudnn.benchmark = True
m = ocrl_b200.SLATE(*slate_config())._module._enc.cuda().eval()
for B in (64, 256):
    x = torch.rand(B, 3, 64, 64, device="cuda")
    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = False
        t_fp32 = timeit(lambda: m(x))
        torch.backends.cudnn.allow_tf32 = True
        t_tf32 = timeit(lambda: m(x))
        xc = x.contiguous(memory_format=torch.channels_last)
        mc = ocrl_b200.SLATE(*slate_config())._module._enc.cuda().eval().to(memory_format=torch.channels_last)
        t_tf32_cl = timeit(lambda: mc(xc))
        mb = ocrl_b200.SLATE(*slate_config())._module._enc.cuda().eval().bfloat16().to(memory_format=torch.channels_last)
        xb = xc.bfloat16()
        t_bf16 = timeit(lambda: mb(xb))
    print(f"B={B}: cuDNN conv stack ms  fp32 {t_fp32:.3f}  tf32 {t_tf32:.3f}  tf32 NHWC {t_tf32_cl:.3f}  bf16 NHWC {t_bf16:.3f}")
