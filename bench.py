#!/usr/bin/env python
"""Benchmark of the slot-attention hot path: SLATE slot-encode images/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--kv fp32|bf16] [--impl reference]

One "step" = one batch of B synthetic 64x64 frames per GPU through ``SLATE.__call__`` (CNN
encoder -> token stage -> fused T-iteration slot-attention kernel), K=6 slots, T=3 iterations,
D=192 (configs/ocr/slate.yaml + README override), seeded random-init weights.

Prints ONE JSON line (rank 0).  ``value`` = images/s with frames resident in HBM; ``e2e`` = the
same through the public API from pinned HOST buffers (H2D of the frames and D2H of the slots in
the timed region); ``roofline`` = fused iteration kernel, algorithmic bytes / CUDA-event time
against the measured HBM peak; ``cpu_baseline`` = the oracle port on the host cores.
``--impl reference`` times the reference's CPU path (oracle port, all host threads) instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "slot-encode images/s"
UNIT = "images/s"
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per step")
    ap.add_argument("--size", type=int, default=64, help="frame size (64 -> N=4096 tokens)")
    ap.add_argument("--slots", type=int, default=6)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--slot-size", type=int, default=192)
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"],
                    help="bf16: bf16 k/v + tensor cores (north-star bf16 mode, 2e-2 parity); fp32: exact-parity mode")
    ap.add_argument("--no-graph", action="store_true", help="do not replay the step from a CUDA graph")
    ap.add_argument("--no-other-mode", action="store_true", help="skip the short run of the other precision mode")
    ap.add_argument("--impl", default="ocrl_b200", choices=["ocrl_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the short SLATE training-step measurement")
    ap.add_argument("--pool", type=int, default=512, help="distinct frames in the synthetic pool")
    ap.add_argument("--sustain", type=float, default=1.2, help="seconds of the extra sustained run (0: skip)")
    ap.add_argument("--sweep", action="store_true",
                    help="BASELINE.json config 5: the iteration kernel over N x K x T (one JSON line per shape, this GPU)")
    return ap.parse_args()


def workload(a):
    return {"workload": f"SLATE encode, {a.size}x{a.size} random-N5C4S4S2 frames (N={a.size * a.size} tokens), "
                        f"K={a.slots}, T={a.iters}, D={a.slot_size}, batch {a.batch}/GPU",
            "frame": a.size, "tokens": a.size * a.size, "num_slots": a.slots, "num_iterations": a.iters,
            "slot_size": a.slot_size, "batch_per_gpu": a.batch,
            "weights": "seeded random-init (pretrained_encoders/slate.pth absent from the reference checkout)"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores (oracle port; the Python reference cannot travel)
# ------------------------------------------------------------------------------------------------
def cpu_encode_fn(a, frames_u8):
    from oracle import slot_oracle as so
    from ocrl_b200.config import slate_config
    import ocrl_b200

    torch.manual_seed(0)
    model = ocrl_b200.SLATE(*slate_config(num_slots=a.slots, num_iterations=a.iters, slot_size=a.slot_size,
                                          mlp_hidden_size=a.slot_size, obs_size=a.size))
    p = {k: v.detach() for k, v in model._module.state_dict().items()}
    obs = frames_u8[: a.batch].permute(0, 3, 1, 2).float() / 255.0
    g = torch.Generator().manual_seed(1)

    def step():
        with torch.no_grad():
            noise = torch.randn(a.batch, a.slots, a.slot_size, generator=g)
            return so.slate_encode(obs, noise, p, a.iters)[0]

    return step


def time_cpu(a, frames_u8, steps, warmup):
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    step = cpu_encode_fn(a, frames_u8)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return a.batch * steps / dt, dt / steps * 1e3, cores


def time_cpu_train(a, frames_u8, batch=8, steps=2):
    """The reference's training step (ocrs/base.py:60-74: get_loss + backward + inf-norm clip + Adam) on the host cores:
    the module wiring of ocrl_b200.SLATE with the slot-attention operators replaced by the oracle port (the CPU leg is
    the one place bench.py may execute oracle/), bounded sample of `batch` frames."""
    import ocrl_b200
    from ocrl_b200 import functional as F, slot_attn
    from ocrl_b200.config import slate_config
    from oracle import slot_oracle as so

    class _OracleFn:
        @staticmethod
        def apply(inputs, slots0, T, epsilon, kv, *params):
            return so.slot_attention(inputs, slots0, dict(zip(F.SA_PARAM_ORDER, params)), T, epsilon)

    saved = (F.SlotAttentionFunction, slot_attn.SlotAttention._check)
    F.SlotAttentionFunction = _OracleFn
    slot_attn.SlotAttention._check = lambda self, inputs, slots, fmap=False: None
    try:
        torch.manual_seed(0)
        model = ocrl_b200.SLATE(*slate_config(num_slots=a.slots, num_iterations=a.iters, slot_size=a.slot_size,
                                              mlp_hidden_size=a.slot_size, obs_size=a.size))
        model.train()
        obs = frames_u8[:batch].permute(0, 3, 1, 2).float() / 255.0
        model.update(obs, None, 1000)
        t0 = time.perf_counter()
        for i in range(steps):
            model.update(obs, None, 1001 + i)
        dt = (time.perf_counter() - t0) / steps
    finally:
        F.SlotAttentionFunction, slot_attn.SlotAttention._check = saved
    return batch / dt, dt * 1e3, batch


def run_reference(a, rank, out=sys.stdout):
    """--impl reference: the reference's CPU path for the same metric and config."""
    if rank != 0:
        return
    from ocrl_b200 import synth

    frames = torch.from_numpy(synth.random_objs_frames(max(a.batch, 64), a.size, seed=0))
    steps = max(1, a.steps)
    ips, ms, cores = time_cpu(a, frames, steps, max(1, min(a.warmup, 2)))
    train = None
    if not a.no_train:
        try:
            tips, tms, tb = time_cpu_train(a, frames)
            train = {"value": tips, "unit": "images/s", "ms_per_step": tms, "cores": cores, "kind": "port",
                     "sample": f"2 steps of SLATE.update on {tb} frames (get_loss + backward + clip + Adam), torch CPU ops "
                               f"with the oracle port of the slot-attention loop, fp32, {cores} threads"}
        except Exception as exc:
            train = {"error": repr(exc)[:300]}
    sample = f"{steps} steps of one {a.batch}-frame batch, oracle port (torch CPU ops, fp32), {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload(a),
            "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "train_step": train}
    print(json.dumps(line), file=out, flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def measure(a, mode, steps, warmup, dev, dist, rank, world, pool_host, pool_dev, want_events=True, sustain_s=0.0,
            pool_u8_host=None):
    """One precision mode: eager loop (kernel events -> roofline), CUDA-graph resident loop (value) and
    CUDA-graph end-to-end loop from pinned host frames (e2e)."""
    import ocrl_b200
    from ocrl_b200 import functional as F
    from ocrl_b200.config import slate_config

    os.environ["OCRL_KV_DTYPE"] = "bf16" if mode == "bf16" else "fp32"
    os.environ.pop("OCRL_CONV_DTYPE", None)  # module default: bf16 convs in bf16 mode, fp32 convs otherwise
    torch.backends.cudnn.allow_tf32 = False  # fp32 mode is fp32 end to end, like the reference's CPU path
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    model = ocrl_b200.SLATE(*slate_config(num_slots=a.slots, num_iterations=a.iters, slot_size=a.slot_size,
                                          mlp_hidden_size=a.slot_size, obs_size=a.size))
    model.to(dev)
    model.eval()
    nb = a.pool // a.batch
    torch.manual_seed(1 + rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n_steps, n_warm, kernel_timers=False, join=None):
        for i in range(n_warm):
            fn(i)
        barrier()
        F.KERNEL_EVENTS = [] if kernel_timers else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_steps):
            fn(n_warm + i)
        if join is not None:
            join()  # the timed stream waits for the copy streams: the last D2H is inside the region
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        ev = F.KERNEL_EVENTS
        F.KERNEL_EVENTS = None
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ev

    def batch_dev(i):
        return pool_dev[(i % nb) * a.batch:(i % nb + 1) * a.batch]

    def batch_host(i):
        return pool_host[(i % nb) * a.batch:(i % nb + 1) * a.batch]

    def step_eager(i):
        with torch.no_grad():
            return model(batch_dev(i))

    ms_eager, events = timed(step_eager, steps, warmup, kernel_timers=False)
    if want_events:
        # Per-kernel CUDA events of the same eager steps.  The kernels of the step are now shorter than the host time of
        # the Python call that launches them, so events recorded in a free-running eager loop bracket host gaps as well;
        # here the stream is held back (a spin kernel of ~20 ms) while the host enqueues every step, and the events then
        # bracket kernels that run back to back in step order (inputs produced by the preceding kernel of the step).
        barrier()
        F.KERNEL_EVENTS = []
        torch.cuda._sleep(int(4e7))
        for i in range(steps):
            step_eager(warmup + i)
        barrier()
        events, F.KERNEL_EVENTS = F.KERNEL_EVENTS, None
    # kernels of libocrl_sa.so per step: the library counts its own launches (ocrl_launch_count); one warm eager step
    # launches what one graph replay holds as kernel nodes (the graphs are captured from the same call)
    from ocrl_b200 import abi
    c0 = abi.lib().ocrl_launch_count()
    step_eager(0)
    torch.cuda.synchronize()
    own_per_step = int(abi.lib().ocrl_launch_count() - c0)

    out_host = torch.empty(a.batch, a.slots, a.slot_size, dtype=torch.float32).pin_memory()
    graphed = None
    if not a.no_graph:
        try:
            graphed = ocrl_b200.GraphedEncoder(model, batch_dev(0))
        except Exception as exc:  # report, fall back to eager launches
            sys.stderr.write(f"bench.py: CUDA graph capture failed ({exc}); timing eager launches\n")
    if graphed is not None:
        def step_single(i):
            graphed(batch_dev(i))

        # the public streaming API: three batches in flight (three graph buffers replaying on their own streams, so the
        # latency-bound iteration kernel of one batch shares the GPU with the convolutions of the next).
        # value: frames already resident in HBM, slots left in HBM.  e2e: every step copies its frames from pinned host
        # memory and its slots back to pinned host memory, copies overlapping the replays.
        nbuf = int(os.environ.get("OCRL_BENCH_BUFFERS", 3))
        icl = os.environ.get("OCRL_BENCH_ITER_CLUSTERS")  # experiment knob (0: launcher's choice); the bench line is the API default
        streamed = ocrl_b200.StreamedEncoder(model, batch_dev(0), buffers=nbuf,
                                             iter_clusters="auto" if icl is None else (int(icl) or None))
        outs_host = [torch.empty_like(out_host).pin_memory() for _ in range(nbuf)]
        outs_dev = [torch.empty(a.batch, a.slots, a.slot_size, device=dev) for _ in range(nbuf)]

        def step_resident(i):
            streamed.submit(batch_dev(i), outs_dev[i % nbuf])

        def step_e2e(i):
            streamed.submit(batch_host(i), outs_host[i % nbuf])
        e2e_join = streamed.join
    else:
        obs_stage = torch.empty(a.batch, 3, a.size, a.size, device=dev)
        step_resident = step_eager
        step_single = None
        e2e_join = None

        def step_e2e(i):
            obs_stage.copy_(batch_host(i), non_blocking=True)
            with torch.no_grad():
                out_host.copy_(model(obs_stage), non_blocking=True)

    ms_res, _ = timed(step_resident, steps, warmup, join=e2e_join)
    ms_single = timed(step_single, steps, warmup)[0] if step_single is not None else None
    ms_e2e, _ = timed(step_e2e, steps, max(3, warmup), join=e2e_join)
    ms_u8 = None
    if graphed is not None and mode == "bf16" and pool_u8_host is not None:
        # the same host-to-host loop fed with the frames as the datasets / environments hold them (uint8 HWC,
        # utils/datasets.py:17): the `/ 255` ingest runs inside the first convolution, a quarter of the H2D bytes
        try:
            streamed_u8 = ocrl_b200.StreamedEncoder(model, pool_u8_host[:a.batch].to(dev), buffers=nbuf)

            def step_u8(i):
                streamed_u8.submit(pool_u8_host[(i % nb) * a.batch:(i % nb + 1) * a.batch], outs_host[i % nbuf])

            ms_u8, _ = timed(step_u8, steps, max(3, warmup), join=streamed_u8.join)
        except Exception as exc:
            sys.stderr.write(f"bench.py: uint8 frame path not measured ({exc})\n")
    rollout = None
    if graphed is not None and mode == "bf16" and want_events:
        # BASELINE.json config 4 (PPO rollout, sb3s/ocr_extractor.py:45): pooling(ocr(obs)) at the rollout batch (4
        # environments) and the PPO minibatch (32), one CUDA-graph replay per call, one call at a time (latency)
        try:
            from types import SimpleNamespace as NS

            pcfg = NS(d_model=128, nhead=8, num_layers=1, pos_emb="None", norm_first=False, use_mlp1=False, use_mlp2=False,
                      cw_embedding=False, push_embedding=False)
            pool = ocrl_b200.Transformer_Module(model.rep_dim, model.num_slots, pcfg).to(dev).eval()
            rollout = {}
            for rb in (4, 32):
                enc = ocrl_b200.GraphedEncoder(ocrl_b200.RolloutExtractor(model, pool), pool_dev[:rb])
                ms_r = timed(lambda i: enc(pool_dev[(i % 8) * rb:(i % 8 + 1) * rb]), 200, 20)[0]
                rollout[f"batch_{rb}"] = {"us_per_call": ms_r / 200 * 1e3, "images_per_s": rb * 200 / ms_r * 1e3 * world}
            rollout["path"] = ("GraphedEncoder(RolloutExtractor(SLATE, Transformer_Module)): frames -> slots -> pooled "
                               "features [B,128], every kernel hand-written, one graph replay per call")
        except Exception as exc:
            sys.stderr.write(f"bench.py: rollout path not measured ({exc})\n")
    sustained = None
    if sustain_s > 0:  # the same resident loop for >= sustain_s seconds, with its own clock samples (rank 0)
        n_sus = max(steps, int(sustain_s * 1e3 / max(ms_res / steps, 1e-3)) + 1)
        clk = Clocks(dev.index if dev.index is not None else 0)
        if rank == 0:
            clk.start()
        ms_sus, _ = timed(step_resident, n_sus, 3, join=e2e_join)
        ms_sus_e2e, _ = timed(step_e2e, n_sus, 3, join=e2e_join)
        sustained = {"value": a.batch * n_sus * world / ms_sus * 1e3, "e2e_value": a.batch * n_sus * world / ms_sus_e2e * 1e3,
                     "unit": UNIT, "steps": n_sus, "seconds": ms_sus / 1e3, "e2e_seconds": ms_sus_e2e / 1e3,
                     "clocks": clk.stop() if rank == 0 else None}
    images = a.batch * steps * world
    res = {"value": images / ms_res * 1e3, "ms_per_step": ms_res / steps, "eager_value": images / ms_eager * 1e3,
           "e2e_value": images / ms_e2e * 1e3, "e2e_ms_per_step": ms_e2e / steps, "graph": graphed is not None,
           "single_stream_value": (images / ms_single * 1e3) if ms_single else None,
           "events": events or [], "own_per_step": own_per_step, "sustained": sustained,
           "e2e_u8_value": (images / ms_u8 * 1e3) if ms_u8 else None, "rollout": rollout}
    return res


def measure_train(a, dev, pool_dev, dist, rank, world, steps=8, warmup=3):
    """SLATE OCR training step (BASELINE.json config 3: num_slots 6, num_iterations 3, bf16 k/v), batch per GPU as in
    the encode metric: ``SLATE.update`` = get_loss + backward + gradient all-reduce (N > 1) + inf-norm clip + Adam.
    The slot-attention loop and the k/v projection run the hand-written forward and backward kernels; the dVAE, the
    transformer decoder, the CNN encoder's backward and the optimizer are torch / cuDNN (library code), so this is a
    context number, not a roofline.  N > 1: data parallel over NCCL (ocrl_b200.dp: bucketed all-reduce launched from
    gradient hooks on a side stream, under the rest of the backward); ``exposed_comm_ms`` is what the compute stream
    still waits for when the backward has finished (CUDA events around the join), max over ranks."""
    import ocrl_b200
    from ocrl_b200 import dp, functional as F
    from ocrl_b200.config import slate_config

    os.environ["OCRL_KV_DTYPE"] = "bf16"
    torch.backends.cudnn.allow_tf32 = True   # torch's GPU defaults, i.e. what the reference's training runs with
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    model = ocrl_b200.SLATE(*slate_config(num_slots=a.slots, num_iterations=a.iters, slot_size=a.slot_size,
                                          mlp_hidden_size=a.slot_size, obs_size=a.size))
    model.to(dev)
    model.train()
    comm_events, reduce_bytes = [], 0
    if world > 1:
        dp.make_data_parallel(model)
        reducer = model._grad_reducer
        reduce_bytes = sum((sum(p.numel() for p in b) + len(b)) * 4 for b in reducer.buckets)
        finish = reducer.finish

        def timed_finish():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            finish()
            e1.record()
            comm_events.append((e0, e1))

        model._after_backward = timed_finish
    nb = a.pool // a.batch

    def step(i):
        return model.update(pool_dev[(i % nb) * a.batch:(i % nb + 1) * a.batch], None, 1000 + i)

    for i in range(warmup):
        step(i)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    comm_events.clear()
    F.KERNEL_EVENTS = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        m = step(warmup + i)
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ev, F.KERNEL_EVENTS = F.KERNEL_EVENTS, None
    ms = e0.elapsed_time(e1) / steps
    comm_ms = sum(x.elapsed_time(y) for x, y in comm_events) / steps if comm_events else 0.0
    if dist is not None:
        t = torch.tensor([ms, comm_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, comm_ms = float(t[0]), float(t[1])
    per = {}
    for name, s0, s1 in ev:
        per.setdefault(name, []).append(s0.elapsed_time(s1))
    return {"value": a.batch * world / ms * 1e3, "unit": "images/s", "ms_per_step": ms, "steps": steps, "n_gpus": world,
            "scaling": "weak", "batch_per_gpu": a.batch,
            "loss": float(m["loss"].detach()), "kernels_ms": {k: sum(v) / len(v) for k, v in per.items()},
            "allreduce_bytes_per_step": reduce_bytes, "exposed_comm_ms": comm_ms,
            "collective": ("NCCL all-reduce of %d gradient buckets per step, launched from gradient hooks on a side stream"
                           % len(model._grad_reducer.buckets)) if world > 1 else None,
            "config": "SLATE.update (get_loss + backward + clip + Adam), bf16 k/v, batch %d per GPU, eager launches; "
                      "slot-attention fwd/bwd + k/v projection hand-written, dVAE / decoder / CNN backward library code "
                      "(torch defaults: cuDNN TF32 on, fp32 matmul)"
                      % a.batch}


def step_mb(a, N, D, esz):
    """MB of device buffers one encode step writes and reads back (bf16 mode: two padded channels-last feature maps,
    x^ [N,64] bf16 (factored form), attention [N,K] fp32; fp32 mode: NCHW maps, k and v)"""
    if a.mode == "bf16":
        maps = 2 * (2 + a.batch * (a.size + 2)) * (a.size + 4) * 64 * 2
        return (maps + a.batch * N * (64 * 2 + a.slots * 4)) / 1e6
    return a.batch * N * (2 * 64 * 4 + 2 * D * esz + a.slots * 4) / 1e6


def roofline_of(a, mode, events):
    it_ms = [s.elapsed_time(e) for name, s, e in events if name == "sa_iter_fwd"]
    tk_ms = [s.elapsed_time(e) for name, s, e in events if name in ("kv_proj_fwd", "xhat_fwd")]
    # bf16 inference runs the FACTORED form where the kernels cover the shape (include/ocrl_sa.h, ocrl_sa_iter_fwd_xhat):
    # the loop streams x^ [N,64] bf16 instead of k, v [N,D] each.  `achieved` keeps SURVEY 8(d)'s algorithmic bytes of the
    # reference formulation (what the contract defines); `streamed_*` says what this kernel actually has to move.
    factored = any(name == "xhat_fwd" for name, _, _ in events)
    N, D, K = a.size * a.size, a.slot_size, a.slots
    esz = 2 if mode == "bf16" else 4
    bytes_img = 2 * N * D * esz + N * K * 4 + 2 * K * D * 4
    streamed_img = N * 64 * 2 + N * K * 4 + 2 * K * D * 4 if factored else bytes_img
    tok_bytes_img = N * 64 * (2 if mode == "bf16" else 4) + 2 * N * D * esz   # reference formulation: tokens in, k and v out
    tok_streamed_img = N * 64 * 2 + N * 64 * 2 if factored else tok_bytes_img   # factored: bf16 map in, x^ out
    peak, peak_src = HBM_FALLBACK_GBS, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        pass
    # DRAM traffic per launch (dram__bytes_read + write) from the `ncu --set full` capture under profiles/r2/ -- used only
    # when that capture was taken from THIS build (digest of the kernel sources) and this configuration, else null:
    # a number taken under a profiler is never timed here, and a capture of another build says nothing about this one
    traffic, tok_traffic = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r2", f"traffic_{mode}.json")) as f:
            tj = json.load(f)
        with open(os.path.join(ROOT, "ocrl_b200", "csrc", "build", "digest.txt")) as f:
            digest = f.read().strip()
        if tj.get("build_digest") == digest and tj.get("config") == [a.batch, N, D, K, a.iters]:
            traffic = tj.get("sa_iter_fwd", {}).get("traffic_bytes")
            tok_traffic = tj.get("kv_proj_fwd", {}).get("traffic_bytes")
    except Exception:
        pass
    it_avg = sum(it_ms) / max(1, len(it_ms))
    tk_avg = sum(tk_ms) / max(1, len(tk_ms))
    achieved = a.batch * bytes_img / (it_avg * 1e-3) / 1e9 if it_ms else None
    tok_achieved = a.batch * tok_bytes_img / (tk_avg * 1e-3) / 1e9 if tk_ms else None
    streamed = a.batch * streamed_img / (it_avg * 1e-3) / 1e9 if it_ms else None
    return {"kernel": (("sa_iter_fwd_umma_kernel, factored form (persistent clusters, one TMA ring of x^ tiles, tcgen05 token pass "
                        "with TMEM accumulators, projections folded into the update weights, three update streams)") if factored else
                       ("sa_iter_fwd_umma_kernel (persistent clusters, TMA ring, tcgen05 token pass with TMEM accumulators, "
                        "two update streams)")) if mode == "bf16" else "sa_iter_fwd_kernel (fp32 FFMA)",
            "form": "factored (streams x^, k = s W_k x^ and v = W_v x^ never materialised)" if factored else "k/v",
            "streamed_bytes_per_image": streamed_img, "streamed_gbs": streamed,
            "frac_streamed": (streamed / peak if streamed else None),
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": (achieved / peak if achieved else None), "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": a.batch * bytes_img,
            "algorithmic_bytes_per_image": bytes_img, "avg_launch_ms": it_avg, "launches_timed": len(it_ms),
            "token_stage": {"kernel": ("kv_proj_tc2_kernel (tcgen05 + TMA, two row groups per CTA" + (", x^ output: stops after norm_inputs)" if factored else ")"))
                            if mode == "bf16" else "token_stage_kernel (fp32 FFMA)",
                            "avg_launch_ms": tk_avg, "achieved": tok_achieved, "unit": "GB/s",
                            "frac": (tok_achieved / peak if tok_achieved else None), "traffic": tok_traffic,
                            "algorithmic_bytes_per_image": tok_bytes_img, "streamed_bytes_per_image": tok_streamed_img,
                            "frac_streamed": (a.batch * tok_streamed_img / (tk_avg * 1e-3) / 1e9 / peak if tk_ms else None)}}


def run_sweep(a, dev, out):
    """BASELINE.json config 5: the fused iteration kernel alone over N in {4096, 16384} x K in {6, 8, 11, 16} x
    T in {3, 5, 7}, bf16 k/v, D = H = 192: CUDA-event time of back-to-back launches on inputs larger than the L2,
    algorithmic bytes per image of SURVEY 8(d), which kernel the dispatcher chose.  The path shards by image with no
    collective, so N GPUs run N copies of each line (weak scaling, see the N-GPU bench lines)."""
    from ocrl_b200 import abi, functional as F
    from oracle import slot_oracle as so

    peak = HBM_FALLBACK_GBS
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
    except Exception:
        pass
    D = a.slot_size
    for N in (4096, 16384):
        B = 64 if N == 4096 else 16
        x = torch.randn(B, N, 64, device=dev)
        for K in (6, 8, 11, 16):
            p = {k: v.to(dev) for k, v in so.random_sa_params(K, 64, D, D, seed=3).items()}
            k, v, _ = F.kv_project(x, p, kv="bf16")
            s0 = torch.randn(B, K, D, device=dev)
            xh, _, _ = F.kv_project(x, p, kv="bf16", xhat_only=True)
            for T in (3, 5, 7):
                bytes_img = 2 * N * D * 2 + N * K * 4 + 2 * K * D * 4
                # both forms of the loop: "factored" (inference: streams x^, what SLATE.__call__ runs) and "k/v" (the
                # training forward).  Ten launches replayed from one CUDA graph: no host time between the kernels
                for form in ("factored", "k/v"):
                    prep = F.PreparedWeights()
                    if form == "factored":
                        fn = lambda: F.iterate_xhat(xh, s0, p, T, prepared=prep)  # noqa: E731
                    else:
                        fn = lambda: F.iterate(k, v, s0, p, T, prepared=prep)  # noqa: E731
                    reps = 10
                    st = torch.cuda.Stream()
                    with torch.cuda.stream(st):
                        for _ in range(3):
                            fn()
                        st.synchronize()
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=st):
                            for _ in range(reps):
                                fn()
                    kern = F.last_kernel()
                    torch.cuda.synchronize()
                    g.replay()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    g.replay()
                    e1.record()
                    torch.cuda.synchronize()
                    us = e0.elapsed_time(e1) / reps * 1e3
                    gbs = B * bytes_img / us / 1e3
                    moved = (N * 64 * 2 if form == "factored" else 2 * N * D * 2)
                    print(json.dumps({"sweep": "iteration kernel", "form": form, "N": N, "K": K, "T": T, "D": D, "B": B, "kernel": kern,
                                      "us": round(us, 1), "images_per_s": round(B / us * 1e6), "GBps": round(gbs, 1),
                                      "frac_of_hbm": round(gbs / peak, 4), "streamed_GBps": round(T * B * moved / us / 1e3, 1),
                                      "algorithmic_bytes_per_image": bytes_img}), file=out, flush=True)
                    del g


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries (NCCL prints its version banner with C stdio)
    write to file descriptor 1 too, so fd 1 is pointed at stderr and the JSON line goes to a private duplicate."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    a = parse()
    out = _claim_stdout()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if a.impl == "reference":
        run_reference(a, rank, out)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the ocrl_b200 path has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    from ocrl_b200 import synth

    if os.environ.get("OCRL_DEV_PROJ_VARIANT"):  # development knob: 1 = the single-chain token-stage kernel
        from ocrl_b200 import abi
        abi.lib().ocrl_dev_proj_variant(int(os.environ["OCRL_DEV_PROJ_VARIANT"]))
    if a.sweep:
        if rank == 0:
            run_sweep(a, dev, out)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    # each rank owns its own shard of the frame pool (no data-path collective: images are independent)
    pool_u8 = torch.from_numpy(synth.random_objs_frames(a.pool, a.size, seed=1000 + rank))
    pool_host = synth.to_obs(pool_u8).contiguous().pin_memory()  # float32 CHW in [0,1], as the reference API takes
    pool_dev = pool_host.to(dev)
    assert a.pool // a.batch >= 1, "--pool must be >= --batch"

    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    main_res = measure(a, a.mode, a.steps, a.warmup, dev, dist, rank, world, pool_host, pool_dev, sustain_s=a.sustain,
                       pool_u8_host=pool_u8.contiguous().pin_memory())
    clk = clocks.stop() if rank == 0 else None
    other = None
    if not a.no_other_mode:
        om = "fp32" if a.mode == "bf16" else "bf16"
        other = (om, measure(a, om, max(5, a.steps // 3), 3, dev, dist, rank, world, pool_host, pool_dev))

    train = None
    if not a.no_train:  # every rank takes part (gradient all-reduce when N > 1)
        try:
            train = measure_train(a, dev, pool_dev, dist, rank, world)
        except Exception as exc:  # the encode metric stands on its own
            train = {"error": repr(exc)[:300]}

    if rank == 0:
        N, D = a.size * a.size, a.slot_size
        esz = 2 if a.mode == "bf16" else 4
        abi_calls = len(main_res["events"]) // max(a.steps, 1)   # kv_proj_fwd + sa_iter_fwd per eager step
        assert abi_calls >= 2 and main_res["own_per_step"] >= 2, "the hand-written kernels did not run in the step"
        line = {"metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None,
                "dtype": ("bf16 (bf16 convs / k / v / tensor-core operands, f32 accumulate, f32 slot update)"
                          if a.mode == "bf16" else "f32"),
                "data": "synthetic",
                "config": workload(a),
                "notes": {"mode": a.mode,
                          "l2": ("inputs larger than L2: a step touches %.0f MB of its own buffers (two padded feature maps, x^ or "
                                 "k/v, attention) and the three graphs in flight own separate buffers (%.0f MB rotating > 126 MB L2); "
                                 "frame pool of %d rotates"
                                 % (step_mb(a, N, D, esz), 3 * step_mb(a, N, D, esz), a.pool)),
                          "kernels": "every kernel of the step is hand-written CUDA in libocrl_sa.so (bf16 mode: mma.sync first "
                                     "convolution, tcgen05 implicit-GEMM 64->64 convolutions with paired taps, tcgen05 token stage "
                                     "(x^ output), tcgen05 iteration kernel in its factored form); torch only draws the slot noise" if a.mode == "bf16" else
                                     "fp32 parity mode: cuDNN convolutions (library), FFMA token stage and iteration kernel",
                          "launch": ("CUDA graph replays of SLATE.__call__, three batches in flight on three streams "
                                     "(single_stream_value: one replay at a time)") if main_res["graph"] else "eager launches",
                          "e2e": "ocrl_b200.StreamedEncoder: per step H2D from pinned frames, graph replay, D2H to pinned "
                                 "slots; copies of neighbouring steps overlap the replays (three buffers)"},
                "clocks": clk,
                "e2e": {"value": main_res["e2e_value"], "unit": UNIT, "ms_per_step": main_res["e2e_ms_per_step"],
                        "h2d_bytes_per_step": a.batch * 3 * a.size * a.size * 4,
                        "d2h_bytes_per_step": a.batch * a.slots * a.slot_size * 4},
                # launches of the library's own kernels inside the timed region of `value`, all ranks
                "gpu_launches": main_res["own_per_step"] * a.steps * world,
                "own_kernels_per_step": main_res["own_per_step"], "eager_value": main_res["eager_value"],
                "single_stream_value": main_res["single_stream_value"],
                "sustained": main_res["sustained"],
                "rollout": main_res.get("rollout"),
                "e2e_u8": ({"value": main_res["e2e_u8_value"], "unit": UNIT, "h2d_bytes_per_step": a.batch * a.size * a.size * 3,
                            "d2h_bytes_per_step": a.batch * a.slots * a.slot_size * 4,
                            "note": "host to host from pinned uint8 HWC frames (the reference's dataset format); `/ 255` inside the first convolution"}
                           if main_res.get("e2e_u8_value") else None),
                "roofline": roofline_of(a, a.mode, main_res["events"])}
        if other is not None:
            om, r = other
            rf = roofline_of(a, om, r["events"])
            line["other_mode"] = {"mode": om, "value": r["value"], "e2e": r["e2e_value"], "unit": UNIT,
                                  "ms_per_step": r["ms_per_step"], "roofline_frac": rf["frac"],
                                  "iter_kernel_ms": rf["avg_launch_ms"], "token_stage_ms": rf["token_stage"]["avg_launch_ms"]}
        if train is not None:
            line["train_step"] = train
        if world == 1 and not a.no_cpu_baseline:
            csteps = 6
            ips, cms, cores = time_cpu(a, pool_u8, csteps, 1)
            line["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{csteps} batches of {a.batch} frames ({cms:.0f} ms each), oracle port "
                                              f"of the reference path with torch CPU ops, fp32"}
        print(json.dumps(line), file=out, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
