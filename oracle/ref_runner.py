"""Drive the reference's own ``GRPOTrainer._compute_loss`` (lifted into ``oracle/_ref`` by ``oracle/build_ref.py``) on
the CPU: the ``--impl reference`` arm and the ``cpu_baseline`` leg of ``bench.py``, and a cross-check of the port in
the tests.  TEST / BASELINE INFRASTRUCTURE ONLY — nothing under ``swh_trl_b200`` may import this module."""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import types

import torch

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "trl_hotpath.py")


def available() -> bool:
    return os.path.exists(_REF)


def load():
    spec = importlib.util.spec_from_file_location("oracle_ref_trl_hotpath", _REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Metrics(dict):
    def __missing__(self, k):
        self[k] = []
        return self[k]


def make_trainer(mod, *, beta, epsilon_low, epsilon_high, delta, loss_type, importance_sampling_level,
                 max_completion_length, temperature, top_entropy_quantile=1.0):
    """An object with exactly the attributes the two lifted methods read (grpo_trainer.py:2058-2175, :1206-1272)."""
    t = object.__new__(mod.GRPOTrainer)
    t.beta, t.epsilon_low, t.epsilon_high = beta, epsilon_low, epsilon_high
    t.loss_type, t.importance_sampling_level = loss_type, importance_sampling_level
    t.top_entropy_quantile, t.max_completion_length, t.temperature = top_entropy_quantile, max_completion_length, temperature
    t.args = types.SimpleNamespace(delta=delta)
    t.accelerator = types.SimpleNamespace(gather=lambda x: x)  # single process
    t._metrics = {"train": _Metrics(), "eval": _Metrics()}
    t.model_kwarg_keys = set()
    t.model = types.SimpleNamespace(training=True)
    return t


class BakedModel:
    """Stands in for the policy: returns pre-computed ``[B, T + 1, V]`` logits (the reference drops the last position,
    grpo_trainer.py:1252)."""

    training = True

    def __init__(self, logits_full):
        self.logits = logits_full

    def __call__(self, **kw):
        return types.SimpleNamespace(logits=self.logits)


def model_logits(completion_logits: torch.Tensor) -> torch.Tensor:
    """``[B, T, V]`` completion logits -> the ``[B, T + 1, V]`` tensor a model with a 1-token prompt would emit."""
    B, T, V = completion_logits.shape
    full = torch.zeros(B, T + 1, V, dtype=completion_logits.dtype)
    full[:, :T] = completion_logits
    return full


def compute_loss(trainer, mod, logits_full, completion_ids, completion_mask, advantages, old=None, ref=None):
    """One call of the reference's ``_compute_loss``; the fork prints tensors inside it (:2127-2128, :2174), which is
    kept (it is the reference's behaviour) but swallowed."""
    B = completion_ids.shape[0]
    inputs = {"prompt_ids": torch.zeros(B, 1, dtype=torch.long), "prompt_mask": torch.ones(B, 1, dtype=torch.long),
              "completion_ids": completion_ids, "completion_mask": completion_mask, "advantages": advantages}
    if old is not None:
        inputs["old_per_token_logps"] = old
    if ref is not None:
        inputs["ref_per_token_logps"] = ref
    with contextlib.redirect_stdout(io.StringIO()):
        return mod.GRPOTrainer._compute_loss(trainer, BakedModel(logits_full), inputs)
