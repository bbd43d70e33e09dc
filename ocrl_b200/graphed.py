"""CUDA-graph replay of the slot-encode path for a fixed batch shape.

PPO rollouts and dataset encoding call ``ocr(obs)`` with the same shape thousands of times
(sb3s/ocr_extractor.py:45); at ~0.5 ms of GPU work per batch the ~20 kernel launches and tensor
allocations of the eager call are a visible fraction of the step.  ``GraphedEncoder`` captures one call
into a CUDA graph (static input / output buffers, graph-safe RNG for the slot-initialisation noise) and
replays it: same kernels, same results, one launch.
"""
from __future__ import annotations

import torch

from . import functional as F


class GraphedEncoder:
    def __init__(self, ocr, example_obs: torch.Tensor, with_masks: bool = False, warmup: int = 3,
                 iter_clusters: int | None = None):
        """iter_clusters: upper bound on the clusters of the iteration kernel in the captured graph (None: the
        launcher's own choice, lowest latency of one replay); passed as ocrl_sa_launch_opts.max_clusters for the
        launches made during the capture (a graph keeps the grid it was captured with)."""
        assert example_obs.is_cuda, "GraphedEncoder captures a CUDA graph"
        self._ocr = ocr
        self._with_masks = with_masks
        self._warmup = warmup
        self._iter_clusters = iter_clusters
        self.static_obs = example_obs.clone()
        self._capture()

    def _params(self):
        mod = getattr(self._ocr, "_module", self._ocr)
        return list(mod.parameters()) if hasattr(mod, "parameters") else []

    def _weights_key(self):
        """The captured graph reads device copies derived from the parameters (bf16 convolution weights, position table,
        exp(log sigma), the iteration kernel's prepared weights): it is valid for exactly this set of parameter versions."""
        self._plist = self._params()
        return tuple((p.data_ptr(), p._version) for p in self._plist)

    def _version_sum(self):
        s = 0
        for p in self._plist:
            s += p._version
        return s

    def _capture(self):
        dev = self.static_obs.device
        opts = lambda: F.launch_options(max_clusters=self._iter_clusters or 0)  # noqa: E731 (single-use context managers)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with opts(), torch.cuda.stream(side), torch.no_grad():
            for _ in range(self._warmup):  # warm up allocator, lazy module state, the weight caches the graph will read
                self._call()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with opts(), torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = self._call()
        self._key = self._weights_key()
        self._vsum = self._version_sum()
        self._calls = 0

    def stale(self) -> bool:
        """True when a parameter changed (optimizer step, load_state_dict, .to()) since the graph was captured.  Every
        in-place update bumps a parameter's version, so the per-call check is the sum of the versions (a few us: at
        the rollout batch the replay itself is ~130 us and host time is what a caller sees); the full key, which also
        notices replaced storages, is compared every 64th call."""
        self._calls += 1
        if self._version_sum() != self._vsum:
            return True
        if self._calls % 64 == 0:
            return self._key != tuple((p.data_ptr(), p._version) for p in self._plist)
        return False

    def recapture(self):
        torch.cuda.synchronize(self.static_obs.device)
        self._capture()

    def _call(self):
        if self._with_masks:
            return self._ocr(self.static_obs, with_masks=True)
        return self._ocr(self.static_obs)

    def __call__(self, obs: torch.Tensor):
        """Copies ``obs`` (host or device) into the static input, replays, returns the static output
        (valid until the next call).  A graph whose parameters changed since the capture is captured again first."""
        if self.stale():
            self.recapture()
        self.static_obs.copy_(obs, non_blocking=True)
        self.graph.replay()
        return self.static_out


class StreamedEncoder:
    """Host-to-host slot encoding at the GPU's pace: ``buffers`` (default three) ``GraphedEncoder`` buffers in
    rotation, the pinned-host -> device copy of batch i+1 and the device -> pinned-host copy of result i-1 run on copy
    streams while the graph of batch i replays (dataset encoding / PPO rollout feeding, utils/datasets.py + sb3s/ocr_extractor.py:45 use case).

        enc = StreamedEncoder(ocr, example_obs_dev)
        for i, frames in enumerate(pinned_batches):
            enc.submit(frames, out_pinned[i])      # asynchronous
        enc.synchronize()                          # every out_pinned[i] is valid

    Every submit performs its own H2D and D2H copy; nothing is cached between batches.

    iter_clusters: clusters of the iteration kernel per replay.  The kernel is latency-bound, so with several replays
    in flight fewer, fuller clusters cost fewer SM-seconds and leave SMs to the neighbours' convolutions: at batch 64
    six clusters (ten or eleven images each) give 165 k images/s (174 k sustained over a second) against 157 k (163 k)
    with the launcher's thirteen, while one replay alone gets slower (profiles/r2/stream_cluster_sweep.txt).
    "auto" (default): one cluster per ten images from batch 48 up when the replays run concurrently, else the launcher's
    choice; None: always the launcher's choice (lowest latency per replay); an integer caps at that count.  Results do
    not depend on it (tested)."""

    def __init__(self, ocr, example_obs: torch.Tensor, with_masks: bool = False, concurrent_replays: bool = True,
                 buffers: int = 3, iter_clusters: int | str | None = "auto"):
        dev = example_obs.device
        self._nb = buffers
        if iter_clusters == "auto":
            batch = example_obs.shape[0]
            iter_clusters = max(1, batch // 10) if (concurrent_replays and buffers > 1 and batch >= 48) else None
        self.iter_clusters = iter_clusters
        self._enc = [GraphedEncoder(ocr, example_obs, with_masks, iter_clusters=iter_clusters) for _ in range(buffers)]
        # with concurrent_replays the buffers replay on their own streams, so the latency-bound iteration kernel of
        # one batch (13 clusters of 8 SMs) shares the GPU with the convolutions of the next
        self._compute = [torch.cuda.Stream(device=dev) for _ in range(buffers)] if concurrent_replays else None
        self._in = torch.cuda.Stream(device=dev)
        self._out = torch.cuda.Stream(device=dev)
        self._dev = dev
        self._ev_in = [torch.cuda.Event() for _ in range(buffers)]
        self._ev_done = [torch.cuda.Event() for _ in range(buffers)]
        self._ev_out = [torch.cuda.Event() for _ in range(buffers)]
        self._used = [False] * buffers
        self._i = 0

    def submit(self, obs_host: torch.Tensor, out_host):
        """obs_host: pinned host (or device) frames of the captured shape; out_host: pinned host tensor (or tuple of
        tensors when with_masks) that receives the result."""
        s = self._i % self._nb
        self._i += 1
        enc = self._enc[s]
        if enc.stale():  # parameters changed since the capture: drain the pipeline, capture every buffer again
            self.synchronize()
            for e in self._enc:
                e.recapture()
        caller = torch.cuda.current_stream(self._dev)
        main = self._compute[s] if self._compute is not None else caller
        if self._compute is not None and not self._used[s]:
            main.wait_stream(caller)                # work the caller queued before the first submit
        if self._used[s]:
            self._in.wait_event(self._ev_done[s])   # the graph that read this input buffer has finished
        if obs_host.is_cuda:
            self._in.wait_stream(caller)            # a device input may still be in production on the caller's stream
        with torch.cuda.stream(self._in):
            enc.static_obs.copy_(obs_host, non_blocking=True)
            self._ev_in[s].record(self._in)
        main.wait_event(self._ev_in[s])
        if self._used[s]:
            main.wait_event(self._ev_out[s])        # the previous result of this buffer has left the device
        with torch.cuda.stream(main):
            enc.graph.replay()
            self._ev_done[s].record(main)
        self._out.wait_event(self._ev_done[s])
        with torch.cuda.stream(self._out):
            if isinstance(enc.static_out, (tuple, list)):
                for dst, src in zip(out_host, enc.static_out):
                    dst.copy_(src, non_blocking=True)
            else:
                out_host.copy_(enc.static_out, non_blocking=True)
            self._ev_out[s].record(self._out)
        self._used[s] = True

    def join(self):
        """Make the current stream wait for every copy submitted so far (no host synchronisation)."""
        main = torch.cuda.current_stream(self._dev)
        main.wait_stream(self._in)
        main.wait_stream(self._out)
        if self._compute is not None:
            for st in self._compute:
                main.wait_stream(st)

    def synchronize(self):
        self._in.synchronize()
        self._out.synchronize()
        if self._compute is not None:
            for st in self._compute:
                st.synchronize()
        torch.cuda.current_stream(self._dev).synchronize()
