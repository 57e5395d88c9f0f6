"""Run the reference's own ``PPOTrainer.train`` / ``RLOOTrainer.train`` (lifted verbatim into ``oracle/_ref`` by
``oracle/build_ref.py``) on a tiny fake trainer: toy policy / value / reward models, a scripted generation step and
stubs for the control plane (accelerator, callbacks, scheduler).  TEST INFRASTRUCTURE ONLY.

Two module instances are made from the same lifted source: one is left as the reference wrote it (its hot functions are
the lifted reference functions), the other goes through ``swh_trl_b200.patch.patch_module`` — leaf functions rebound,
``train`` rewritten by ``train_patch``.  Same seeds, same initial weights: the logged metrics and the trained
parameters of the two runs must agree.
"""
from __future__ import annotations

import contextlib
import copy
import gc
import importlib.util
import math
import os
import time
import types
from collections import defaultdict

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
LOOPS = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "trl_train_loops.py")
HOTPATH = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "trl_hotpath.py")

PAD, EOS = 0, 1


def available() -> bool:
    return os.path.exists(LOOPS) and os.path.exists(HOTPATH)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class ToyLM(nn.Module):
    """Per-token language model: logits[b, l] depend on token l and on the mean embedding of the prefix."""

    def __init__(self, vocab, hidden, out):
        super().__init__()
        self.emb = nn.Embedding(vocab, hidden)
        self.mix = nn.Linear(hidden, hidden)
        self.head = nn.Linear(hidden, out)

    def forward(self, ids):
        e = self.emb(ids)
        prefix = e.cumsum(1) / torch.arange(1, ids.shape[1] + 1, device=ids.device).view(1, -1, 1)
        return self.head(torch.tanh(self.mix(e) + prefix))


class PolicyAndValue(nn.Module):
    def __init__(self, policy, value_model):
        super().__init__()
        self.policy, self.value_model = policy, value_model


def _forward(model, query_responses, pad_token_id):  # stands in for trl.trainer.utils.forward
    if isinstance(model, PolicyAndValue):
        return types.SimpleNamespace(logits=model.policy(query_responses)), model.value_model(query_responses)
    return types.SimpleNamespace(logits=model(query_responses))


def _get_reward(model, query_responses, pad_token_id, context_length):  # trl.trainer.utils.get_reward
    full = model(query_responses)                                      # [B, L, 1]
    not_pad = query_responses[:, context_length:] != pad_token_id
    last = context_length - 1 + not_pad.long().sum(1).clamp(min=1)
    score = full.squeeze(-1)[torch.arange(full.shape[0], device=full.device), last]
    return full, score, last


class _Gen:
    """Scripted sampling: the same responses in every run (their randomness is not the trainer's)."""

    def __init__(self, seed, response_length, temperature, vocab):
        self.seed, self.T, self.temp, self.V = seed, response_length, temperature, vocab
        self.calls = 0

    def __call__(self, policy, queries, local_bs, pad_token_id, generation_config):
        g = torch.Generator().manual_seed(self.seed + self.calls)
        self.calls += 1
        B = queries.shape[0]
        resp = torch.randint(2, self.V, (B, self.T), generator=g)
        stop = torch.randint(self.T // 2, self.T + 2, (B,), generator=g)  # some rows never emit EOS
        for b in range(B):
            if stop[b] < self.T:
                resp[b, stop[b]] = EOS
        qr = torch.cat([queries, resp.to(queries.device)], 1)
        pol = policy.policy if isinstance(policy, PolicyAndValue) else policy
        logits = pol(qr)[:, queries.shape[1] - 1:-1] / (self.temp + 1e-7)
        return qr, logits


def _namespace(hot):
    ns = dict(torch=torch, nn=nn, np=np, math=math, time=time, gc=gc, defaultdict=defaultdict,
              GenerationConfig=lambda **kw: types.SimpleNamespace(**kw), INVALID_LOGPROB=1.0,
              empty_cache=lambda: None, forward=_forward, get_reward=_get_reward,
              selective_log_softmax=hot.selective_log_softmax, masked_mean=hot.masked_mean,
              masked_whiten=hot.masked_whiten, first_true_indices=hot.first_true_indices,
              truncate_response=hot.truncate_response)

    @contextlib.contextmanager
    def unwrap_model_for_generation(model, accelerator, gather_deepspeed3_params=True):
        yield model
    ns["unwrap_model_for_generation"] = unwrap_model_for_generation
    return ns


def load_loops(tag: str):
    """A fresh module holding the two lifted ``train`` methods with their globals injected."""
    hot = _load(HOTPATH, f"ref_hotpath_{tag}")
    mod = _load(LOOPS, f"ref_train_loops_{tag}")
    for k, v in _namespace(hot).items():
        setattr(mod, k, v)
    return mod


class _Accelerator:
    def __init__(self, device, ga):
        self.device, self.gradient_accumulation_steps = device, ga

    def print(self, *a, **k):
        pass

    def unwrap_model(self, m):
        return m

    @contextlib.contextmanager
    def accumulate(self, model):
        yield

    def backward(self, loss):  # accelerate scales by 1 / gradient_accumulation_steps
        (loss / self.gradient_accumulation_steps).backward()

    def gather_for_metrics(self, x):
        return x

    def gather(self, x):
        return x


class _Callbacks:
    def on_train_begin(self, args, state, control):
        return control

    on_step_end = on_save = on_train_end = on_train_begin


def make_trainer(mod, kind, device, seed=0, vocab=96, hidden=16, ctx=5, T=8, **over):
    torch.manual_seed(seed)
    policy, value_model = ToyLM(vocab, hidden, vocab), ToyLM(vocab, hidden, 1)
    ref_policy, reward_model = copy.deepcopy(policy), ToyLM(vocab, hidden, 1)
    with torch.no_grad():  # the policy has drifted a little from the reference policy: kl != 0
        for p in policy.parameters():
            p.add_(0.05 * torch.randn_like(p))
    k = 2 if kind == "rloo" else 1
    n_prompts = 4
    B = n_prompts * k
    args = types.SimpleNamespace(
        num_total_batches=2, batch_size=B, local_batch_size=B, local_mini_batch_size=B // 2, num_mini_batches=2,
        per_device_train_batch_size=B // 4, gradient_accumulation_steps=2, num_ppo_epochs=2, response_length=T,
        temperature=0.7, local_rollout_forward_batch_size=B // 2, ds3_gather_for_generation=False,
        missing_eos_penalty=1.0, kl_coef=0.05, kl_estimator="k1", whiten_rewards=False, gamma=1.0, lam=0.95,
        cliprange=0.2, cliprange_value=0.2, vf_coef=0.1, logging_steps=1, eval_steps=None, save_steps=None,
        num_sample_generations=0, total_episodes=2 * B, stop_token_id=EOS, rloo_k=k, normalize_reward=False,
        reward_clip_range=10.0, normalize_advantage=False, token_level_kl=True)
    for name, v in over.items():
        setattr(args, name, v)
    g = torch.Generator().manual_seed(seed + 7)
    data = [{"input_ids": torch.randint(2, vocab, (n_prompts, ctx), generator=g)} for _ in range(2)]
    model = PolicyAndValue(policy, value_model).to(device) if kind == "ppo" else policy.to(device)
    cls = mod.PPOTrainer if kind == "ppo" else mod.RLOOTrainer
    t = object.__new__(cls)
    t.args, t.accelerator = args, _Accelerator(torch.device(device), args.gradient_accumulation_steps)
    t.model, t.ref_model, t.reward_model = model, ref_policy.to(device), reward_model.to(device)
    t.ref_policy = t.ref_model  # RLOOTrainer's name for it (rloo_trainer.py:250)
    t.optimizer = torch.optim.Adam(model.parameters(), lr=3e-3)
    t.processing_class = types.SimpleNamespace(pad_token_id=PAD, eos_token_id=EOS)
    t.dataloader, t.train_dataset_len = data, n_prompts * 2
    t.state = types.SimpleNamespace(global_step=0, episode=0)
    t.callback_handler, t.control = _Callbacks(), types.SimpleNamespace(should_save=False)
    t.lr_scheduler = types.SimpleNamespace(step=lambda: None, get_last_lr=lambda: [3e-3])
    t.is_deepspeed_enabled, t.stop_token_id, t.sample_generations_freq = False, EOS, 1
    t.logged = []
    t.log = lambda metrics: t.logged.append(dict(metrics))
    mod.batch_generation = _Gen(seed + 100, T, args.temperature, vocab)
    return t


def run(mod, kind, device, seed=0, **over):
    """``(logged metrics per update, trained parameters)`` of ``train()`` as ``mod`` defines it."""
    t = make_trainer(mod, kind, device, seed, **over)
    np.random.seed(seed)
    torch.manual_seed(seed + 1)
    t.train()
    params = {n: p.detach().float().cpu().clone() for n, p in t.model.named_parameters()}
    return t.logged, params
