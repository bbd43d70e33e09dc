"""Drop-in OCR wrappers: ``Base`` (reference: ocrs/base.py:8-88) and ``SLATE``
(ocrs/slate/slate.py:13-69).  Same attributes (``name``, ``rep_dim``, ``num_slots``, ``_module``,
``_opt``), same methods, same checkpoint dictionary."""
import math

import torch
from torch.nn.utils import clip_grad_norm_

from .adjacent import linear_warmup
from .slate_module import SLATE_Module


def optimizer_to(optim, device):
    for state in optim.state.values():
        for key, val in (state.items() if isinstance(state, dict) else []):
            if isinstance(val, torch.Tensor):
                state[key] = val.to(device)


class Base:
    def __init__(self, ocr_config, env_config) -> None:
        self.name = ocr_config.name
        self._config = ocr_config
        self._obs_size = env_config.obs_size
        self._obs_channels = env_config.obs_channels
        self.rep_dim = self._module.rep_dim
        self.num_slots = self._module.num_slots
        learning = getattr(self._config, "learning", None)
        if learning is not None and hasattr(learning, "lr"):
            self._opt = torch.optim.Adam(self._module.parameters(), lr=learning.lr)

    def __call__(self, obs):
        return self._module(obs)

    def wandb_watch(self, config):
        import wandb

        wandb.watch(self._module, log=config.log)

    def get_loss(self, obs, with_rep=False) -> dict:
        return self._module.get_loss(obs, with_rep)

    def train(self) -> None:
        self._module.train()

    def eval(self) -> None:
        self._module.eval()

    def to(self, device) -> None:
        self._module.to(device)
        if hasattr(self, "_opt"):
            optimizer_to(self._opt, device)

    def set_zero_grad(self):
        if hasattr(self, "_opt"):
            self._opt.zero_grad()

    def do_step(self):
        if hasattr(self, "_opt"):
            self._opt.step()

    def get_samples(self, obs) -> dict:
        return self._module.get_samples(obs)

    def _after_backward(self):
        """Hook between backward and the gradient clip; data-parallel wrappers reduce gradients here."""

    def update(self, obs, masks, step: int) -> dict:
        if not hasattr(self, "_opt"):
            return {}
        self._opt.zero_grad()
        metrics = self.get_loss(obs, masks)
        metrics["loss"].backward()
        self._after_backward()
        learning = self._config.learning
        if hasattr(learning, "clip"):
            norm_type = learning.clip_norm_type if hasattr(learning, "clip_norm_type") else "inf"
            metrics["norm"] = clip_grad_norm_(self._module.parameters(), learning.clip, norm_type)
        self._opt.step()
        return metrics

    def save(self) -> dict:
        ckpt = {"ocr_module_state_dict": self._module.state_dict()}
        if hasattr(self, "_opt"):
            ckpt["ocr_opt_state_dict"] = self._opt.state_dict()
        return ckpt

    def load(self, checkpoint) -> None:
        self._module.load_state_dict(checkpoint["ocr_module_state_dict"])
        if hasattr(self, "_opt"):
            self._opt.load_state_dict(checkpoint["ocr_opt_state_dict"])


class SLATE(Base):
    def __init__(self, ocr_config, env_config) -> None:
        self._module = SLATE_Module(ocr_config, env_config)
        super().__init__(ocr_config, env_config)
        lr = self._config.learning
        self._opt = torch.optim.Adam([
            {"params": self._module.get_dvae_params(), "lr": lr.lr_dvae},
            {"params": self._module.get_sa_params(), "lr": lr.lr_enc},
            {"params": self._module.get_tfdec_params(), "lr": lr.lr_dec},
        ])

    def __call__(self, obs, with_attns=False, with_masks=False):
        return self._module(obs, with_attns, with_masks)

    def get_loss(self, obs, masks, with_rep=False, with_mse=False) -> dict:
        out = self._module.get_loss(obs, masks, with_rep, with_mse)
        metrics, rep = out if with_rep else (out, None)
        groups = self._opt.param_groups
        metrics.update({"lr_dvae": torch.Tensor([groups[0]["lr"]]), "lr_enc": torch.Tensor([groups[1]["lr"]]),
                        "lr_dec": torch.Tensor([groups[2]["lr"]])})
        return (metrics, rep) if with_rep else metrics

    def update(self, obs, masks, step: int) -> dict:
        self._module.update_tau(step)
        lr = self._config.learning
        warm = linear_warmup(step, 0, 1, 0, lr.lr_warmup_steps)
        decay = math.exp(step / lr.lr_half_life * math.log(0.5))
        self._opt.param_groups[0]["lr"] = lr.lr_dvae
        self._opt.param_groups[1]["lr"] = decay * warm * lr.lr_enc
        self._opt.param_groups[2]["lr"] = decay * warm * lr.lr_dec
        return super().update(obs, masks, step)
