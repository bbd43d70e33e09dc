"""CUDA-graph replay of the slot-encode path for a fixed batch shape.

PPO rollouts and dataset encoding call ``ocr(obs)`` with the same shape thousands of times
(sb3s/ocr_extractor.py:45); at ~0.5 ms of GPU work per batch the ~20 kernel launches and tensor
allocations of the eager call are a visible fraction of the step.  ``GraphedEncoder`` captures one call
into a CUDA graph (static input / output buffers, graph-safe RNG for the slot-initialisation noise) and
replays it: same kernels, same results, one launch.
"""
from __future__ import annotations

import torch


class GraphedEncoder:
    def __init__(self, ocr, example_obs: torch.Tensor, with_masks: bool = False, warmup: int = 3):
        assert example_obs.is_cuda, "GraphedEncoder captures a CUDA graph"
        self._ocr = ocr
        self._with_masks = with_masks
        self.static_obs = example_obs.clone()
        side = torch.cuda.Stream(device=example_obs.device)
        side.wait_stream(torch.cuda.current_stream(example_obs.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):  # warm up allocator, cuDNN autotuner, lazy module state
                self._call()
        torch.cuda.current_stream(example_obs.device).wait_stream(side)
        torch.cuda.synchronize(example_obs.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = self._call()

    def _call(self):
        if self._with_masks:
            return self._ocr(self.static_obs, with_masks=True)
        return self._ocr(self.static_obs)

    def __call__(self, obs: torch.Tensor):
        """Copies ``obs`` (host or device) into the static input, replays, returns the static output
        (valid until the next call)."""
        self.static_obs.copy_(obs, non_blocking=True)
        self.graph.replay()
        return self.static_out
