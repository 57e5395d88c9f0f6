#!/usr/bin/env python
"""Which summation order does torch use for ``rewards.view(-1, G).mean(dim=1)`` / ``.std(dim=1)`` (grpo_trainer.py:1921-1922)?
Compares torch (CPU and, if present, CUDA) bit for bit with candidate orders computed in numpy: fp32 sequential, fp32
pairwise tree, fp32 four strided lanes, and the double-accumulated value K3 produces.  Prints the agreeing fraction."""
import json, sys
import numpy as np
import torch


def candidates(xs):
    n, G = xs.shape
    out = {}
    acc = np.zeros(n, dtype=np.float32)
    for i in range(G):
        acc = (acc + xs[:, i]).astype(np.float32)
    out["seq_f32"] = acc / np.float32(G)
    out["double"] = (xs.astype(np.float64).sum(1) / G).astype(np.float32)
    if G & (G - 1) == 0:
        t = xs.copy()
        while t.shape[1] > 1:
            t = (t[:, 0::2] + t[:, 1::2]).astype(np.float32)
        out["tree_f32"] = t[:, 0] / np.float32(G)
    if G % 4 == 0:
        lanes = np.zeros((n, 4), dtype=np.float32)
        for i in range(0, G, 4):
            lanes = (lanes + xs[:, i:i + 4]).astype(np.float32)
        out["lanes4_seq"] = (((lanes[:, 0] + lanes[:, 1]).astype(np.float32) + lanes[:, 2]).astype(np.float32) + lanes[:, 3]).astype(np.float32) / np.float32(G)
        out["lanes4_tree"] = ((lanes[:, 0] + lanes[:, 1]).astype(np.float32) + (lanes[:, 2] + lanes[:, 3]).astype(np.float32)).astype(np.float32) / np.float32(G)
    return out


def std_candidates(xs):
    n, G = xs.shape
    d = xs.astype(np.float64)
    out = {"double_two_pass": np.sqrt(((d - d.mean(1, keepdims=True)) ** 2).sum(1) / (G - 1)).astype(np.float32)}
    mean = np.zeros(n, dtype=np.float32)
    m2 = np.zeros(n, dtype=np.float32)
    for i in range(G):  # Welford in fp32
        delta = (xs[:, i] - mean).astype(np.float32)
        mean = (mean + delta / np.float32(i + 1)).astype(np.float32)
        m2 = (m2 + delta * (xs[:, i] - mean).astype(np.float32)).astype(np.float32)
    out["welford_f32_seq"] = np.sqrt(m2 / np.float32(G - 1)).astype(np.float32)
    return out


res = {}
devs = ["cpu"] + (["cuda"] if torch.cuda.is_available() else [])
for G in (2, 4, 8, 16):
    g = torch.Generator().manual_seed(G)
    x = torch.randn(200000, G, generator=g) * 3
    xs = x.numpy()
    for dev in devs:
        m = x.to(dev).mean(dim=1).cpu().numpy()
        s = x.to(dev).std(dim=1).cpu().numpy()
        res[f"{dev} G={G} mean"] = {k: float((v == m).mean()) for k, v in candidates(xs).items()}
        res[f"{dev} G={G} std"] = {k: float((v == s).mean()) for k, v in std_candidates(xs).items()}
print(json.dumps(res, indent=1))
