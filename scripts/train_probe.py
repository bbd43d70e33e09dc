"""Training-step and encode timings of the BASELINE.json configurations on one GPU (context numbers next to the
bench line): `update()` = get_loss + backward + inf-norm clip + Adam, batch 64, 64x64 synthetic frames.

  C2 small  Slot-Attention (broadcast decoder), K=6, T=7, D=64 / H=128
  C2 large  Slot-Attention (broadcast decoder), K=6, T=7, D=192
  C3        SLATE, K=6, T=3, D=192

each with fp32 k/v (parity mode) and bf16 k/v; per step the time spent in the hand-written kernels is listed.
Usage: python scripts/train_probe.py [out.json]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocrl_b200  # noqa: E402
from ocrl_b200 import functional as F  # noqa: E402
from ocrl_b200 import synth  # noqa: E402
from ocrl_b200.config import slate_config, slot_attention_config  # noqa: E402


def run(name, cfg, kv, frames, B=64, steps=6, warmup=3):
    os.environ["OCRL_KV_DTYPE"] = kv
    torch.manual_seed(0)
    model = ocrl_b200.SLATE(*cfg)
    model.to("cuda")
    model.train()
    nb = frames.shape[0] // B

    def step(i):
        return model.update(frames[(i % nb) * B:(i % nb + 1) * B], None, 1000 + i)

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    F.KERNEL_EVENTS = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        m = step(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    kern = {}
    for nme, a, b in F.KERNEL_EVENTS:
        kern[nme] = kern.get(nme, 0.0) + a.elapsed_time(b) / steps
    F.KERNEL_EVENTS = None
    # encode (inference) of the same model
    model.eval()
    with torch.no_grad():
        for i in range(3):
            model(frames[:B])
        torch.cuda.synchronize()
        e0.record()
        for i in range(10):
            model(frames[(i % nb) * B:(i % nb + 1) * B])
        e1.record()
        torch.cuda.synchronize()
    enc_ms = e0.elapsed_time(e1) / 10
    rec = {"config": name, "kv": kv, "train_ms_per_step": round(ms, 3), "train_images_per_s": round(B / ms * 1e3, 1),
           "loss": float(m["loss"]), "kernels_ms": {k: round(v, 4) for k, v in kern.items()},
           "encode_eager_ms": round(enc_ms, 3), "encode_eager_images_per_s": round(B / enc_ms * 1e3, 1)}
    print(json.dumps(rec), flush=True)
    del model
    torch.cuda.empty_cache()
    return rec


if __name__ == "__main__":
    torch.backends.cudnn.benchmark = True
    frames = synth.to_obs(torch.from_numpy(synth.random_objs_frames(256, 64, seed=0))).contiguous().cuda()
    out = []
    for name, cfg in (("C2 small (Slot-Attention, D=64, H=128, T=7, broadcast decoder)", slot_attention_config(False)),
                      ("C2 large (Slot-Attention, D=192, T=7, broadcast decoder)", slot_attention_config(True)),
                      ("C3 (SLATE, D=192, T=3)", slate_config())):
        for kv in ("fp32", "bf16"):
            out.append(run(name, cfg, kv, frames))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
