"""Tensor-level wrappers over the C-ABI: argument marshalling only (pointers, sizes, the current stream).

PyTorch is plumbing here — it owns device memory and the stream; every numeric result below is produced by a
kernel in ``libb200trl.so``.  All wrappers are asynchronous and never synchronise.
"""

from __future__ import annotations

import contextlib
import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import GrpoCfg, check, lib

_DTYPES = {torch.bfloat16: _lib.DTYPE_BF16, torch.float16: _lib.DTYPE_F16, torch.float32: _lib.DTYPE_F32,
           torch.float64: _lib.DTYPE_F64}

launch_count = 0  # kernels launched through this module (bench.py reports it as gpu_launches)


def _count(n: int = 1) -> None:
    global launch_count
    launch_count += n


def _ptr(t: Optional[torch.Tensor]):
    # a plain int: ctypes converts it for a c_void_p parameter itself (every prototype in _lib.PROTOTYPES is typed)
    return None if t is None else t.data_ptr()


def _raw_stream(device: torch.device) -> int:
    """The current stream's ``cudaStream_t`` of ``device`` as an int: the raw getter costs ~0.3 us where
    ``torch.cuda.current_stream(device).cuda_stream`` costs ~8 us (it builds a Stream object) -- three of those per step
    were a fifth of a config-1 step's host time."""
    idx = device.index
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device() if idx is None else idx)


def _stream(t: torch.Tensor):
    return _raw_stream(t.device)


def _need_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: swh_trl_b200 has no CPU path (got device {t.device})")


def _f32(t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    _need_cuda(t, name)
    if t.dtype is torch.float32 and t.is_contiguous():
        return t.detach() if t.requires_grad else t
    return t.detach().to(torch.float32).contiguous()


def _as(t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """``t.to(dtype).contiguous()`` without the two no-op dispatcher trips when ``t`` already is both."""
    return t if (t.dtype is dtype and t.is_contiguous()) else t.to(dtype).contiguous()


class _Keep:
    """Owns the fp32 / contiguous temporaries of one C call.

    ``_ptr(_f32(x))`` alone would hand the library a pointer whose tensor CPython frees before the call is made:
    two same-sized temporaries in one argument list then alias in the caching allocator (e.g. bf16 ``old`` and
    ``ref`` log-probs).  Every converted tensor is therefore parked here until the wrapper returns; after that the
    allocator's stream ordering keeps the block valid for the launched kernel.
    """

    __slots__ = ("held",)

    def __init__(self):
        self.held = []

    def f32(self, t: Optional[torch.Tensor], name: str):
        x = _f32(t, name)
        if x is not None:
            self.held.append(x)
        return _ptr(x)


def collapse_rows(shape, strides) -> Optional[int]:
    """Row stride (in elements) if a ``(..., V)`` layout is addressable as ``base + r * row_stride``, else None.

    Pure shape arithmetic (unit-tested on CPU): the last dim must be dense and the leading dims (size-1 dims
    ignored) must form one arithmetic progression.
    """
    V = shape[-1]
    if V != 1 and strides[-1] != 1:
        return None
    dims = [(s, st) for s, st in zip(shape[:-1], strides[:-1]) if s != 1]
    if any(dims[i][1] != dims[i + 1][0] * dims[i + 1][1] for i in range(len(dims) - 1)):
        return None
    row_stride = dims[-1][1] if dims else V
    return row_stride if row_stride >= V else None


def collapse_rows2(shape, strides):
    """``(row_stride, rows_per_batch, batch_stride)`` for a ``(..., V)`` layout addressable with at most two
    levels (batch, row) — e.g. the ``logits[:, :-1][:, -T:]`` view of a ``[B, L, V]`` model output — else None.
    ``rows_per_batch == 0`` means flat.  Pure shape arithmetic (unit-tested on CPU)."""
    V = shape[-1]
    if V != 1 and strides[-1] != 1:
        return None
    groups = []
    for s, st in zip(shape[:-1], strides[:-1]):
        if s == 1:
            continue
        if groups and groups[-1][1] == s * st:
            groups[-1] = (groups[-1][0] * s, st)
        else:
            groups.append((s, st))
    if not groups:
        return V, 0, 0
    if len(groups) == 1:
        return (groups[0][1], 0, 0) if groups[0][1] >= V else None
    if len(groups) == 2:
        (_, bs), (rpb, rs) = groups
        if rs >= V and bs >= (rpb - 1) * rs + V:
            return rs, rpb, bs
    return None


class Rows:
    """A logits-like tensor resolved to the C-ABI's row addressing."""

    __slots__ = ("t", "n", "V", "row_stride", "rows_per_batch", "batch_stride")

    def __init__(self, t, n, V, row_stride, rows_per_batch=0, batch_stride=0):
        self.t, self.n, self.V = t, n, V
        self.row_stride, self.rows_per_batch, self.batch_stride = row_stride, rows_per_batch, batch_stride


def rows_view(logits: torch.Tensor, rows_per_batch: Optional[int] = None) -> Rows:
    """Resolve ``logits (..., V)`` to row addressing without copying when its layout has at most two levels.

    ``rows_per_batch``: if given, a two-level layout must split exactly there (the fused kernels index per-sequence
    data by ``row // T``).  Otherwise one contiguous copy is made — the copy the reference always makes at
    grpo_trainer.py:1252-1258.
    """
    _need_cuda(logits, "logits")
    if logits.dtype not in _DTYPES:
        raise TypeError(f"unsupported logits dtype {logits.dtype}")
    if logits.dim() < 1:
        raise ValueError("logits must have a vocabulary dimension")
    V = logits.shape[-1]
    n = logits.numel() // V if V else 0
    lay = collapse_rows2(tuple(logits.shape), tuple(logits.stride())) if n else (V, 0, 0)
    if lay is not None:
        rs, rpb, bs = lay
        if rpb == 0 or rows_per_batch is None or rpb == rows_per_batch:
            return Rows(logits, n, V, rs, rpb, bs)
    return Rows(logits.contiguous(), n, V, V)


def alloc_dlogits(r: Rows, shape):
    """``(dlogits, dl_row_stride, dl_batch_stride)`` for logits addressed as ``r``.

    Normally a contiguous tensor.  When the bf16 rows do not start on 16-byte boundaries (vocab % 8 != 0 as in GPT-2's
    50 257, or a misaligned view) the resident kernel needs every dlogits row at the same offset inside its 16-byte
    granule as the logits row it mirrors (its interior vectors are then aligned stores and only the two edge vectors of
    a row go out element by element), so the buffer copies the logits' strides modulo 8 elements: at most 7 elements
    of padding per row / per batch, returned as a strided view of the requested shape."""
    x, V = r.t, r.V
    skew = x.dtype in (torch.bfloat16, torch.float16) and bool(V % 8 or r.row_stride % 8 or r.batch_stride % 8 or x.data_ptr() % 16)
    if not skew or len(shape) not in (2, 3) or r.n == 0:
        return torch.empty(shape, dtype=x.dtype, device=x.device), V, 0
    e0 = (x.data_ptr() // 2) % 8
    rs = V + ((r.row_stride - V) % 8)
    if r.rows_per_batch:
        nb = r.n // r.rows_per_batch
        bs_min = r.rows_per_batch * rs
        bs = bs_min + ((r.batch_stride - bs_min) % 8)
        buf = torch.empty(e0 + nb * bs + 8, dtype=x.dtype, device=x.device)
        return buf.as_strided(tuple(shape), (bs, rs, 1), e0), rs, bs
    buf = torch.empty(e0 + r.n * rs + 8, dtype=x.dtype, device=x.device)
    strides = (rs, 1) if len(shape) == 2 else (shape[1] * rs, rs, 1)
    return buf.as_strided(tuple(shape), strides, e0), rs, 0


def make_cfg(beta: float, epsilon_low: float, epsilon_high: float, delta: Optional[float], loss_type: str,
             importance_sampling_level: str, max_completion_length: int, grad_scale: float = 1.0,
             skip_masked: bool = False) -> GrpoCfg:
    """Pack GRPO hyper-parameters; unknown enums raise the reference's ValueErrors (grpo_trainer.py:2106-2109, 2137)."""
    if loss_type not in _lib.LOSS_TYPES:
        raise ValueError(f"Unknown loss type: {loss_type}")
    if importance_sampling_level not in _lib.IS_LEVELS:
        raise ValueError(
            f"Unknown importance sampling level: {importance_sampling_level}. Possible values are 'token' "
            "and 'sequence'.")
    cfg = GrpoCfg()
    cfg.beta = float(beta)
    cfg.clip_low = 1.0 - float(epsilon_low)  # double arithmetic, rounded once to fp32 like torch's scalar clamp
    cfg.clip_high = 1.0 + float(epsilon_high)
    cfg.delta = float(delta) if delta is not None else 0.0
    cfg.has_delta = 0 if delta is None else 1
    cfg.loss_type = _lib.LOSS_TYPES[loss_type]
    cfg.is_level = _lib.IS_LEVELS[importance_sampling_level]
    cfg.max_completion_length = float(max_completion_length)
    cfg.grad_scale = float(grad_scale)
    cfg.skip_masked = int(bool(skip_masked))
    return cfg


# ------------------------------------------------------------------------------------------------ K1
def logprob_entropy_fwd(logits: torch.Tensor, ids: torch.Tensor, inv_temperature: float = 1.0,
                        want_entropy: bool = True, want_lse: bool = True):
    """fp32 ``(logp, entropy|None, lse|None)`` shaped like ``ids`` — one pass over the logits."""
    r = rows_view(logits)
    x, n, V = r.t, r.n, r.V
    _need_cuda(ids, "index")
    idx = _as(ids, torch.int64)
    if idx.numel() != n:
        raise ValueError(f"index has {idx.numel()} elements, logits has {n} rows")
    shape = tuple(logits.shape[:-1])
    logp = torch.empty(shape, dtype=torch.float32, device=x.device)
    ent = torch.empty(shape, dtype=torch.float32, device=x.device) if want_entropy else None
    lse = torch.empty(shape, dtype=torch.float32, device=x.device) if want_lse else None
    if n:
        check(lib.b200trl_logprob_entropy_fwd(_ptr(x), _DTYPES[x.dtype], n, V, r.row_stride, r.rows_per_batch,
                                              r.batch_stride, _ptr(idx), float(inv_temperature), _ptr(logp),
                                              _ptr(ent), _ptr(lse), _stream(x)), "logprob_entropy_fwd")
        _count()
    return logp, ent, lse


def masked_logprob_fwd(logits: torch.Tensor, ids: torch.Tensor, row_mask: torch.Tensor, inv_temperature: float = 1.0,
                       want_entropy: bool = False, want_lse: bool = True):
    """Like ``logprob_entropy_fwd`` but rows with ``row_mask == False`` are not read (outputs 0) —
    ``b200trl_masked_logprob_fwd``."""
    r = rows_view(logits)
    x, n, V = r.t, r.n, r.V
    _need_cuda(ids, "index")
    idx = _as(ids, torch.int64)
    m = row_mask.to(torch.bool).contiguous().view(torch.uint8)
    if idx.numel() != n or m.numel() != n:
        raise ValueError(f"index / mask have {idx.numel()} / {m.numel()} elements, logits has {n} rows")
    shape = tuple(logits.shape[:-1])
    logp = torch.empty(shape, dtype=torch.float32, device=x.device)
    ent = torch.empty(shape, dtype=torch.float32, device=x.device) if want_entropy else None
    lse = torch.empty(shape, dtype=torch.float32, device=x.device) if want_lse else None
    if n:
        check(lib.b200trl_masked_logprob_fwd(_ptr(x), _DTYPES[x.dtype], n, V, r.row_stride, r.rows_per_batch,
                                             r.batch_stride, _ptr(idx), _ptr(m), float(inv_temperature), _ptr(logp),
                                             _ptr(ent), _ptr(lse), _stream(x)), "masked_logprob_fwd")
        _count()
    return logp, ent, lse


def logprob_bwd(logits: torch.Tensor, ids: torch.Tensor, lse: torch.Tensor, g: torch.Tensor,
                inv_temperature: float = 1.0) -> torch.Tensor:
    """``dlogits`` (logits dtype, contiguous) for per-token upstream gradient ``g``."""
    k = _Keep()
    r = rows_view(logits)
    x, n, V = r.t, r.n, r.V
    idx = _as(ids, torch.int64)
    out, dl_rs, dl_bs = alloc_dlogits(r, tuple(logits.shape))
    if n:
        check(lib.b200trl_logprob_bwd(_ptr(x), _DTYPES[x.dtype], n, V, r.row_stride, r.rows_per_batch, r.batch_stride,
                                      _ptr(idx), float(inv_temperature), k.f32(lse, "lse"), k.f32(g, "g"),
                                      _ptr(out), dl_rs, dl_bs, _stream(x)), "logprob_bwd")
        _count()
    return out


def mask_stats(mask: torch.Tensor):
    """``(mask_i32, row_count[B], total[1])`` for a ``[B,T]`` completion mask."""
    _need_cuda(mask, "completion_mask")
    m = _as(mask, torch.int32)
    B, T = m.shape
    row = torch.empty(B, dtype=torch.float32, device=m.device)
    tot = torch.empty(1, dtype=torch.float32, device=m.device)
    check(lib.b200trl_mask_stats(_ptr(m), B, T, _ptr(row), _ptr(tot), _stream(m)), "mask_stats")
    _count(2)  # memset node + kernel
    return m, row, tot


def grpo_fused_fwd_bwd(logits, ids, mask_i32, row_count, total_count, advantages, old_logp, ref_logp, cfg: GrpoCfg,
                       inv_temperature: float, want_grad: bool = True, dlogits_out: Optional[torch.Tensor] = None):
    """One pass: ``(logp, entropy, lse, dlogits|None)``; see ``b200trl_grpo_fused_fwd_bwd``."""
    k = _Keep()
    B, T = mask_i32.shape
    r = rows_view(logits, rows_per_batch=T)
    x, n, V = r.t, r.n, r.V
    if n != B * T:
        raise ValueError(f"logits rows {n} != B*T {B * T}")
    idx = _as(ids, torch.int64)
    logp = torch.empty(B, T, dtype=torch.float32, device=x.device)
    ent = torch.empty(B, T, dtype=torch.float32, device=x.device)
    lse = torch.empty(B, T, dtype=torch.float32, device=x.device)
    dl, dl_rs, dl_bs = None, V, 0
    if want_grad:
        dl = dlogits_out if dlogits_out is not None else alloc_dlogits(r, (B, T, V))[0]
        d = collapse_rows2(tuple(dl.shape), tuple(dl.stride()))
        if d is None or (d[1] not in (0, T)) or dl.dtype != x.dtype:
            raise ValueError("dlogits_out must be a [B,T,V] view with at most (batch, row) strides and the logits dtype")
        dl_rs, dl_bs = d[0], d[2]
    check(lib.b200trl_grpo_fused_fwd_bwd(
        _ptr(x), _DTYPES[x.dtype], B, T, V, r.row_stride, r.batch_stride, _ptr(idx), _ptr(mask_i32),
        k.f32(advantages, "advantages"), k.f32(old_logp, "old_per_token_logps"),
        k.f32(ref_logp, "ref_per_token_logps"), C.byref(cfg), float(inv_temperature), _ptr(row_count),
        _ptr(total_count), _ptr(logp), _ptr(ent), _ptr(lse), _ptr(dl), dl_rs, dl_bs, _stream(x)),
        "grpo_fused_fwd_bwd")
    _count()
    return logp, ent, lse, dl


def grpo_fused_step(logits, ids, mask_i32, row_count, total_count, advantages, old_logp, ref_logp, cfg: GrpoCfg,
                    inv_temperature: float, want_grad: bool = True, dlogits_out: Optional[torch.Tensor] = None):
    """``(logp, entropy, lse, dlogits|None, loss[1], metrics[8])`` — ``b200trl_grpo_fused_step``: the fused pass with the
    loss value and the logged metric means produced by the same launch (resident kernel with a gradient; otherwise K2
    runs right behind the pass inside the C call).  ``cfg.grad_scale`` scales ``dlogits`` only, never the loss value.
    ``row_count`` / ``total_count`` may both be ``None``: the call counts ``mask_i32`` itself (no ``mask_stats`` launch)."""
    k = _Keep()
    B, T = mask_i32.shape
    r = rows_view(logits, rows_per_batch=T)
    x, n, V = r.t, r.n, r.V
    if n != B * T:
        raise ValueError(f"logits rows {n} != B*T {B * T}")
    idx = _as(ids, torch.int64)
    dev = x.device
    logp = torch.empty(B, T, dtype=torch.float32, device=dev)
    ent = torch.empty(B, T, dtype=torch.float32, device=dev)
    lse = torch.empty(B, T, dtype=torch.float32, device=dev)
    out = torch.empty(1 + _lib.NUM_GRPO_METRICS, dtype=torch.float32, device=dev)  # loss | metrics, one allocation
    dl, dl_rs, dl_bs = None, V, 0
    if want_grad:
        dl = dlogits_out if dlogits_out is not None else alloc_dlogits(r, (B, T, V))[0]
        d = collapse_rows2(tuple(dl.shape), tuple(dl.stride()))
        if d is None or (d[1] not in (0, T)) or dl.dtype != x.dtype:
            raise ValueError("dlogits_out must be a [B,T,V] view with at most (batch, row) strides and the logits dtype")
        dl_rs, dl_bs = d[0], d[2]
    ws = _workspace(dev, lib.b200trl_grpo_fused_step_workspace_bytes(B), "grpo_fused_step", zero=True)
    check(lib.b200trl_grpo_fused_step(
        _ptr(x), _DTYPES[x.dtype], B, T, V, r.row_stride, r.batch_stride, _ptr(idx), _ptr(mask_i32),
        k.f32(advantages, "advantages"), k.f32(old_logp, "old_per_token_logps"),
        k.f32(ref_logp, "ref_per_token_logps"), C.byref(cfg), float(inv_temperature), _ptr(row_count),
        _ptr(total_count), _ptr(logp), _ptr(ent), _ptr(lse), _ptr(dl), dl_rs, dl_bs, _ptr(ws), _ptr(out),
        out.data_ptr() + 4, _stream(x)), "grpo_fused_step")
    _count()
    return logp, ent, lse, dl, out[:1], out[1:]


# ------------------------------------------------------------------------------------------------ K2
_ws_cache = {}
# a private {(device, family): tensor} while a GraphedStep warms up / captures.  A plain global, not a thread-local:
# autograd runs the backward operators on its own device thread, and they must see the same scope (stream capture is
# process-wide state anyway -- no other thread may issue CUDA work during it)
_ws_private = None


@contextlib.contextmanager
def private_workspaces(store: dict):
    """Scratch buffers requested inside the block live in ``store`` instead of the per-stream cache.

    A captured CUDA graph bakes its workspace POINTERS in: were they the shared per-stream buffers, a later eager call
    that needs a larger one would replace (and free) the buffer under the graph, and a second graph captured on a
    recycled pool stream could run concurrently on the same scratch memory.  ``graphs.GraphedStep`` therefore owns
    its workspaces for as long as the graph lives."""
    global _ws_private
    prev, _ws_private = _ws_private, store
    try:
        yield store
    finally:
        _ws_private = prev


def _workspace(device, nbytes: int, key: str, zero: bool) -> torch.Tensor:
    # one workspace per (device, stream, kernel family): kernels of different streams never share scratch memory
    store = _ws_private
    if store is not None:
        cache, k = store, (device, key)
    else:
        cache, k = _ws_cache, (device, _raw_stream(device), key)
    ws = cache.get(k)
    if ws is None or ws.numel() < nbytes:
        ws = (torch.zeros if zero else torch.empty)(max(nbytes, 256), dtype=torch.uint8, device=device)
        cache[k] = ws
    return ws


def grpo_loss(logp, old_logp, ref_logp, advantages, mask_i32, row_count, total_count, cfg: GrpoCfg,
              ent_mask: Optional[torch.Tensor] = None, entropy: Optional[torch.Tensor] = None, want_g: bool = False):
    """``(loss[1], metrics[8], g[B,T]|None)`` — ``b200trl_grpo_loss``."""
    k = _Keep()
    B, T = mask_i32.shape
    dev = mask_i32.device
    ws = _workspace(dev, lib.b200trl_grpo_loss_workspace_bytes(B), "grpo_loss", zero=True)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    metrics = torch.empty(_lib.NUM_GRPO_METRICS, dtype=torch.float32, device=dev)
    g = torch.empty(B, T, dtype=torch.float32, device=dev) if want_g else None
    em = None if ent_mask is None else ent_mask.to(torch.uint8).contiguous()
    check(lib.b200trl_grpo_loss(
        k.f32(logp, "per_token_logps"), k.f32(old_logp, "old_per_token_logps"),
        k.f32(ref_logp, "ref_per_token_logps"), k.f32(advantages, "advantages"), _ptr(mask_i32), _ptr(em),
        k.f32(entropy, "entropies"), B, T, C.byref(cfg), _ptr(row_count), _ptr(total_count), _ptr(ws), _ptr(loss),
        _ptr(metrics), _ptr(g), _stream(mask_i32)), "grpo_loss")
    _count()
    return loss, metrics, g


def entropy_quantile_mask(entropies: torch.Tensor, mask: torch.Tensor, threshold: float):
    """``(bool mask, threshold[1])`` — ``b200trl_entropy_quantile_mask``."""
    _need_cuda(entropies, "entropies")
    e = _f32(entropies, "entropies")
    m = _as(mask, torch.int32)
    out = torch.empty(e.shape, dtype=torch.bool, device=e.device)  # the kernel writes 0 / 1 bytes: a bool's storage
    thr = torch.empty(1, dtype=torch.float32, device=e.device)
    if e.numel():
        ws = _workspace(e.device, lib.b200trl_entropy_quantile_workspace_bytes(e.numel()), "quantile", zero=False)
        check(lib.b200trl_entropy_quantile_mask(_ptr(e), _ptr(m), e.numel(), float(threshold), _ptr(ws), _ptr(out),
                                                _ptr(thr), _stream(e)), "entropy_quantile_mask")
        _count()
    else:
        out.zero_()
    return out, thr


# ------------------------------------------------------------------------------------------------ K3
def group_advantages(rewards_per_func: torch.Tensor, weights: torch.Tensor, num_generations: int,
                     scale_rewards: bool, local_offset: int, local_count: int):
    """``dict(advantages, all, rewards, mean, std, is_std_zero)`` — ``b200trl_group_advantages``."""
    r = _f32(rewards_per_func, "rewards_per_func")
    if r.dim() == 1:
        r = r.unsqueeze(1)
    w = _f32(weights, "reward_weights")
    Bg, F = r.shape
    if Bg % num_generations:
        raise ValueError(f"global batch {Bg} is not divisible by num_generations {num_generations}")
    dev = r.device
    rewards = torch.empty(Bg, dtype=torch.float32, device=dev)
    adv_all = torch.empty(Bg, dtype=torch.float32, device=dev)
    adv_loc = torch.empty(local_count, dtype=torch.float32, device=dev)
    ng = Bg // num_generations
    mean = torch.empty(ng, dtype=torch.float32, device=dev)
    std = torch.empty(ng, dtype=torch.float32, device=dev)
    zero = torch.empty(ng, dtype=torch.bool, device=dev)  # the kernel writes 0 / 1 bytes: a bool's storage, no cast kernel
    check(lib.b200trl_group_advantages(_ptr(r), _ptr(w), Bg, F, num_generations, int(bool(scale_rewards)),
                                       int(local_offset), int(local_count), _ptr(rewards), _ptr(adv_all),
                                       _ptr(adv_loc), _ptr(mean), _ptr(std), _ptr(zero), _stream(r)),
          "group_advantages")
    _count()
    return dict(advantages=adv_loc, all=adv_all, rewards=rewards, mean=mean, std=std, is_std_zero=zero)


def generation_stats(packed: torch.Tensor, world: int, b_local: int, rewards_per_func: torch.Tensor,
                     mean_grouped: torch.Tensor, std_grouped: torch.Tensor, is_std_zero: torch.Tensor) -> torch.Tensor:
    """fp64 ``[12 + 2 * n_funcs]`` on the device — ``b200trl_generation_stats`` (layout in include/b200trl.h)."""
    k = _Keep()
    _need_cuda(packed, "packed")
    pk = packed.to(torch.int64).contiguous()
    if pk.numel() != world * (1 + 2 * b_local):
        raise ValueError(f"packed has {pk.numel()} elements, expected world * (1 + 2 * B_local) = {world * (1 + 2 * b_local)}")
    r = _f32(rewards_per_func, "rewards_per_func")
    if r.dim() == 1:
        r = r.unsqueeze(1)
    if r.shape[0] != world * b_local:
        raise ValueError(f"rewards_per_func has {r.shape[0]} rows, expected world * B_local = {world * b_local}")
    z = is_std_zero.to(torch.bool).contiguous()
    out = torch.empty(12 + 2 * r.shape[1], dtype=torch.float64, device=pk.device)
    check(lib.b200trl_generation_stats(_ptr(pk), int(world), int(b_local), _ptr(r), int(r.shape[1]),
                                       k.f32(mean_grouped, "mean"), k.f32(std_grouped, "std"), _ptr(z),
                                       int(mean_grouped.numel()), _ptr(out), _stream(pk)), "generation_stats")
    _count()
    return out


# ------------------------------------------------------------------------------------------------ K4 / PPO
def ppo_rewards_gae(logprobs, ref_logprobs, values, scores, sequence_lengths, kl_coef, kl_estimator, gamma, lam,
                    whiten_rewards, want_filled: bool = True):
    """``dict(rewards, advantages, returns[, logprobs, ref_logprobs, values])`` — ``b200trl_ppo_rewards_gae``."""
    if kl_estimator not in _lib.KL_ESTIMATORS:
        raise ValueError(f"Unknown kl_estimator: {kl_estimator}")
    lp, rlp, val = _f32(logprobs, "logprobs"), _f32(ref_logprobs, "ref_logprobs"), _f32(values, "values")
    sc = _f32(scores, "scores")
    sl = _as(sequence_lengths, torch.int64)
    B, T = lp.shape
    dev = lp.device
    ws = _workspace(dev, lib.b200trl_ppo_gae_workspace_bytes(B, T), "ppo_gae", zero=False)
    mk = lambda: torch.empty(B, T, dtype=torch.float32, device=dev)  # noqa: E731
    rewards, adv, ret = mk(), mk(), mk()
    lpf, rlpf, vf = (mk(), mk(), mk()) if want_filled else (None, None, None)
    check(lib.b200trl_ppo_rewards_gae(_ptr(lp), _ptr(rlp), _ptr(val), _ptr(sc), _ptr(sl), B, T, float(kl_coef),
                                      _lib.KL_ESTIMATORS[kl_estimator], float(gamma), float(lam),
                                      int(bool(whiten_rewards)), _ptr(ws), _ptr(rewards), _ptr(adv), _ptr(ret),
                                      _ptr(lpf), _ptr(rlpf), _ptr(vf), _stream(lp)), "ppo_rewards_gae")
    _count()
    out = dict(rewards=rewards, advantages=adv, returns=ret)
    if want_filled:
        out.update(logprobs=lpf, ref_logprobs=rlpf, values=vf)
    return out


def ppo_fused_fwd_bwd(logits, responses, sequence_lengths, old_logprobs, advantages, inv_temperature, cliprange,
                      grad_scale: float = 1.0, want_grad: bool = True):
    """``(new_logprobs, entropy, lse, dlogits|None)`` — ``b200trl_ppo_fused_fwd_bwd``."""
    k = _Keep()
    mb, T = responses.shape
    r = rows_view(logits, rows_per_batch=T)
    x, n, V = r.t, r.n, r.V
    if n != mb * T:
        raise ValueError(f"logits rows {n} != mb*T {mb * T}")
    idx = _as(responses, torch.int64)
    sl = _as(sequence_lengths, torch.int64)
    dev = x.device
    nlp = torch.empty(mb, T, dtype=torch.float32, device=dev)
    ent = torch.empty(mb, T, dtype=torch.float32, device=dev)
    lse = torch.empty(mb, T, dtype=torch.float32, device=dev)
    dl, dl_rs, dl_bs = alloc_dlogits(r, (mb, T, V)) if want_grad else (None, V, 0)
    check(lib.b200trl_ppo_fused_fwd_bwd(_ptr(x), _DTYPES[x.dtype], mb, T, V, r.row_stride, r.batch_stride, _ptr(idx),
                                        _ptr(sl), k.f32(old_logprobs, "old_logprobs"),
                                        k.f32(advantages, "advantages"), float(inv_temperature), float(cliprange),
                                        float(grad_scale), _ptr(nlp), _ptr(ent), _ptr(lse), _ptr(dl), dl_rs, dl_bs,
                                        _stream(x)), "ppo_fused_fwd_bwd")
    _count()
    return nlp, ent, lse, dl


def ppo_fused_step(logits, responses, sequence_lengths, old_logprobs, advantages, returns, values, vpred,
                   inv_temperature, cliprange, cliprange_value, vf_coef, grad_scale: float = 1.0, want_grad: bool = True,
                   want_dvpred: bool = True):
    """``(new_logprobs, entropy, dlogits|None, stats[8], dvpred|None)`` — ``b200trl_ppo_fused_step``: the fused PPO pass
    with the clipped losses, their statistics and ``d loss / d vpred`` produced by the same launch (resident kernel
    with a gradient; otherwise K2p runs right behind the pass inside the C call)."""
    k = _Keep()
    mb, T = responses.shape
    r = rows_view(logits, rows_per_batch=T)
    x, n, V = r.t, r.n, r.V
    if n != mb * T:
        raise ValueError(f"logits rows {n} != mb*T {mb * T}")
    idx = _as(responses, torch.int64)
    sl = _as(sequence_lengths, torch.int64)
    dev = x.device
    nlp = torch.empty(mb, T, dtype=torch.float32, device=dev)
    ent = torch.empty(mb, T, dtype=torch.float32, device=dev)
    dl, dl_rs, dl_bs = alloc_dlogits(r, (mb, T, V)) if want_grad else (None, V, 0)
    dvp = torch.empty(mb, T, dtype=torch.float32, device=dev) if want_dvpred else None
    stats = torch.empty(_lib.NUM_PPO_STATS, dtype=torch.float32, device=dev)
    ws = _workspace(dev, lib.b200trl_ppo_fused_step_workspace_bytes(mb), "ppo_fused_step", zero=True)
    check(lib.b200trl_ppo_fused_step(_ptr(x), _DTYPES[x.dtype], mb, T, V, r.row_stride, r.batch_stride, _ptr(idx), _ptr(sl),
                                     k.f32(old_logprobs, "old_logprobs"), k.f32(advantages, "advantages"),
                                     k.f32(returns, "returns"), k.f32(values, "values"), k.f32(vpred, "vpred"),
                                     float(inv_temperature), float(cliprange), float(cliprange_value), float(vf_coef),
                                     float(grad_scale), _ptr(nlp), _ptr(ent), None, _ptr(dl), dl_rs, dl_bs, _ptr(dvp),
                                     _ptr(ws), _ptr(stats), _stream(x)), "ppo_fused_step")
    _count()
    return nlp, ent, dl, stats, dvp


def ppo_loss(new_logprobs, old_logprobs, advantages, returns, values, vpred, entropy, sequence_lengths, cliprange,
             cliprange_value, vf_coef, grad_scale: float = 1.0, want_dvpred: bool = True):
    """``(stats[8], dvpred|None)`` — ``b200trl_ppo_loss``."""
    k = _Keep()
    nlp = _f32(new_logprobs, "new_logprobs")
    mb, T = nlp.shape
    dev = nlp.device
    sl = _as(sequence_lengths, torch.int64)
    ws = _workspace(dev, lib.b200trl_grpo_loss_workspace_bytes(mb), "ppo_loss", zero=True)
    stats = torch.empty(_lib.NUM_PPO_STATS, dtype=torch.float32, device=dev)
    dvp = torch.empty(mb, T, dtype=torch.float32, device=dev) if want_dvpred else None
    check(lib.b200trl_ppo_loss(_ptr(nlp), k.f32(old_logprobs, "old_logprobs"), k.f32(advantages, "advantages"),
                               k.f32(returns, "returns"), k.f32(values, "values"), k.f32(vpred, "vpred"),
                               k.f32(entropy, "entropy"), _ptr(sl), mb, T, float(cliprange),
                               float(cliprange_value), float(vf_coef), float(grad_scale), _ptr(ws), _ptr(stats),
                               _ptr(dvp), _stream(nlp)), "ppo_loss")
    _count()
    return stats, dvp


# ------------------------------------------------------------------------------------------------ a-12
def masked_whiten(values: torch.Tensor, mask: torch.Tensor, shift_mean: bool = True, want_out: bool = True):
    """``(whitened|None, stats[3]={mean, unbiased var, count})`` — ``b200trl_masked_whiten``."""
    v = _f32(values, "values")
    m = mask.expand_as(values).to(torch.uint8).contiguous()
    dev = v.device
    ws = _workspace(dev, lib.b200trl_masked_workspace_bytes(v.numel()), "masked", zero=False)
    out = torch.empty_like(v) if want_out else None
    stats = torch.empty(3, dtype=torch.float32, device=dev)
    check(lib.b200trl_masked_whiten(_ptr(v), _ptr(m), v.numel(), int(bool(shift_mean)), _ptr(ws), _ptr(out),
                                    _ptr(stats), _stream(v)), "masked_whiten")
    _count()
    return out, stats


def rescale_if_needed(buf: torch.Tensor, actual: torch.Tensor, expected: float) -> None:
    """``buf *= actual / expected`` on the device, skipped when equal (no host sync)."""
    lay = collapse_rows2(tuple(buf.shape), tuple(buf.stride()))
    if lay is None:
        raise ValueError("rescale_if_needed needs a buffer with at most (batch, row) strides")
    row_stride, rpb, bs = lay
    V = buf.shape[-1]
    a = actual.detach()
    if a.dtype is not torch.float32:
        a = a.to(torch.float32)
    check(lib.b200trl_rescale_if_needed_batched(_ptr(buf), _DTYPES[buf.dtype], buf.numel() // V, V, row_stride, rpb, bs,
                                                _ptr(a), float(expected), _stream(buf)), "rescale_if_needed")
    _count()


# ------------------------------------------------------------------------------------------------ RLOO
def rloo_rewards_advantages(logprobs, ref_logprobs, scores, sequence_lengths, kl_coef, rloo_k, normalize_reward,
                            reward_clip_range, normalize_advantage, token_level_kl, want_filled: bool = True):
    """``dict(advantages, rlhf_reward, non_score_reward[, logprobs, ref_logprobs])`` — rloo_trainer.py:397-441."""
    lp, rlp = _f32(logprobs, "logprobs"), _f32(ref_logprobs, "ref_logprobs")
    sc = _f32(scores, "scores")
    sl = _as(sequence_lengths, torch.int64)
    B, T = lp.shape
    dev = lp.device
    adv = torch.empty(B, dtype=torch.float32, device=dev)
    rlhf = torch.empty(B, dtype=torch.float32, device=dev)
    ns = torch.empty(B, dtype=torch.float32, device=dev)
    lpf = torch.empty_like(lp) if want_filled else None
    rlpf = torch.empty_like(rlp) if want_filled else None
    check(lib.b200trl_rloo_rewards_advantages(_ptr(lp), _ptr(rlp), _ptr(sc), _ptr(sl), B, T, float(kl_coef),
                                              int(rloo_k), int(bool(normalize_reward)), float(reward_clip_range),
                                              int(bool(normalize_advantage)), int(bool(token_level_kl)), _ptr(adv),
                                              _ptr(rlhf), _ptr(ns), _ptr(lpf), _ptr(rlpf), _stream(lp)),
          "rloo_rewards_advantages")
    _count()
    out = dict(advantages=adv, rlhf_reward=rlhf, non_score_reward=ns)
    if want_filled:
        out.update(logprobs=lpf, ref_logprobs=rlpf)
    return out


def rloo_loss(new_logprobs, old_logprobs, advantages, entropy, sequence_lengths, cliprange, grad_scale: float = 1.0,
              want_g: bool = True):
    """``(stats[8], g[mb,T]|None)`` — ``b200trl_rloo_loss``."""
    k = _Keep()
    nlp = _f32(new_logprobs, "new_logprobs")
    mb, T = nlp.shape
    dev = nlp.device
    sl = _as(sequence_lengths, torch.int64)
    ws = _workspace(dev, lib.b200trl_grpo_loss_workspace_bytes(mb), "rloo_loss", zero=True)
    stats = torch.empty(8, dtype=torch.float32, device=dev)
    g = torch.empty(mb, T, dtype=torch.float32, device=dev) if want_g else None
    check(lib.b200trl_rloo_loss(_ptr(nlp), k.f32(old_logprobs, "old_logprobs"),
                                k.f32(advantages, "advantages"), k.f32(entropy, "entropy"), _ptr(sl), mb, T,
                                float(cliprange), float(grad_scale), _ptr(ws), _ptr(stats), _ptr(g), _stream(nlp)),
          "rloo_loss")
    _count()
    return stats, g


# ------------------------------------------------------------------------------------------------ K5 (forward)
def fused_linear_logprob_fwd(hidden: torch.Tensor, weight: torch.Tensor, ids: torch.Tensor, inv_temperature: float = 1.0,
                             want_entropy: bool = True, want_lse: bool = False):
    """``(logp, entropy|None, lse|None)`` of ``(hidden @ weight.T) * inv_temperature`` — logits never materialised."""
    _need_cuda(hidden, "hidden")
    if hidden.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise TypeError("fused_linear_logprob_fwd needs bf16 hidden states and weights")
    H = hidden.shape[-1]
    h2 = hidden.reshape(-1, H)
    if h2.stride(-1) != 1:
        h2 = h2.contiguous()
    w = weight if weight.stride(-1) == 1 else weight.contiguous()
    n, V = h2.shape[0], w.shape[0]
    idx = _as(ids, torch.int64)
    if idx.numel() != n:
        raise ValueError(f"index has {idx.numel()} elements, hidden has {n} rows")
    shape = tuple(hidden.shape[:-1])
    dev = hidden.device
    logp = torch.empty(shape, dtype=torch.float32, device=dev)
    ent = torch.empty(shape, dtype=torch.float32, device=dev) if want_entropy else None
    lse = torch.empty(shape, dtype=torch.float32, device=dev) if want_lse else None
    if n:
        ws = _workspace(dev, lib.b200trl_fused_linear_workspace_bytes(n, V), "fused_linear", zero=False)
        check(lib.b200trl_fused_linear_logprob_fwd(_ptr(h2), h2.stride(0), _ptr(w), w.stride(0), n, H, V, _ptr(idx),
                                                   float(inv_temperature), _ptr(ws), _ptr(logp), _ptr(ent), _ptr(lse),
                                                   _stream(h2)), "fused_linear_logprob_fwd")
        _count(2)
    return logp, ent, lse


# ------------------------------------------------------------------------------------------------ a-13: the seam
def fused_linear_grpo(hidden: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], ids: torch.Tensor,
                      mask_i32: torch.Tensor, advantages: torch.Tensor, old_logp: Optional[torch.Tensor],
                      ref_logp: Optional[torch.Tensor], cfg: GrpoCfg, inv_temperature: float, chunk_seqs: int,
                      need_dh: bool, need_dw: bool, need_db: bool, seq_rows=None):
    """``(loss[1], metrics[8], logp, entropy, dH|None, dW bf16|None, db fp32|None)`` — ``b200trl_fused_linear_grpo``:
    the whole chunked lm_head + GRPO loss forward/backward in one C call (GEMMs + K1 in place + K2).  ``dW`` is
    accumulated over the chunks in fp32 and handed back rounded once to the weight's dtype.  ``seq_rows`` (a HOST
    sequence of B ints: index of each sequence's last unmasked token + 1) selects ``b200trl_fused_linear_grpo_trimmed``:
    the rows behind it take part in none of the contractions."""
    k = _Keep()
    _need_cuda(hidden, "_input")
    if hidden.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise TypeError("fused_linear_grpo needs bf16 hidden states and weights")
    B, T, H = hidden.shape
    V = weight.shape[0]
    dev = hidden.device
    h = hidden.contiguous()
    w = weight.contiguous()
    b = None if bias is None else bias.to(torch.bfloat16).contiguous()
    idx = _as(ids, torch.int64)
    logp = torch.empty(B, T, dtype=torch.float32, device=dev)
    ent = torch.empty(B, T, dtype=torch.float32, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    metrics = torch.empty(_lib.NUM_GRPO_METRICS, dtype=torch.float32, device=dev)
    dh = torch.empty_like(h) if need_dh else None
    n_chunks = -(-B // max(1, min(int(chunk_seqs), B)))
    if seq_rows is not None:
        seq_rows = [int(v) for v in seq_rows]
        if len(seq_rows) != B:
            raise ValueError(f"seq_rows has {len(seq_rows)} entries for {B} sequences")
        n_chunks = max(1, sum(1 for v in seq_rows if v > 0))
    dw = torch.empty(V, H, dtype=torch.bfloat16, device=dev) if need_dw else None
    dw_acc = torch.empty(V, H, dtype=torch.float32, device=dev) if need_dw and n_chunks > 1 else None
    db = torch.empty(V, dtype=torch.float32, device=dev) if need_db else None
    ws = _workspace(dev, lib.b200trl_fused_linear_grpo_workspace_bytes(B, T, H, V, int(chunk_seqs)), "fused_linear_grpo",
                    zero=False)
    if seq_rows is not None:
        rows_host = (C.c_int64 * B)(*seq_rows)
        check(lib.b200trl_fused_linear_grpo_trimmed(
            _ptr(h), _ptr(w), _ptr(b), B, T, H, V, _ptr(idx), _ptr(mask_i32), k.f32(advantages, "advantages"),
            k.f32(old_logp, "old_per_token_logps"), k.f32(ref_logp, "ref_per_token_logps"), C.byref(cfg),
            float(inv_temperature), rows_host, _ptr(ws), _ptr(logp), _ptr(ent), _ptr(loss), _ptr(metrics), _ptr(dh),
            _ptr(dw_acc), _ptr(dw), _ptr(db), _stream(h)), "fused_linear_grpo_trimmed")
    else:
        check(lib.b200trl_fused_linear_grpo(
            _ptr(h), _ptr(w), _ptr(b), B, T, H, V, _ptr(idx), _ptr(mask_i32), k.f32(advantages, "advantages"),
            k.f32(old_logp, "old_per_token_logps"), k.f32(ref_logp, "ref_per_token_logps"), C.byref(cfg),
            float(inv_temperature), int(chunk_seqs), _ptr(ws), _ptr(logp), _ptr(ent), _ptr(loss), _ptr(metrics), _ptr(dh),
            _ptr(dw_acc), _ptr(dw), _ptr(db), _stream(h)), "fused_linear_grpo")
    _count(2 + n_chunks + 1)  # mask stats (memset + kernel), K1 per chunk, K2; the GEMMs are library launches
    return loss, metrics, logp, ent, dh, dw, db


# ------------------------------------------------------------------------------------------------ K7: tcgen05 GEMMs
TC_OUT_BF16, TC_OUT_F32_ACC, TC_OUT_F32 = 1, 2, 3


def tc_gemm(a: torch.Tensor, b: torch.Tensor, a_layout: int = 0, b_layout: int = 0,
            out: Optional[torch.Tensor] = None, accumulate: bool = False, bias: Optional[torch.Tensor] = None,
            m_fastest: bool = True, split_k: bool = True, addend: Optional[torch.Tensor] = None,
            out_fp32: bool = False) -> torch.Tensor:
    """``D = A @ B.T`` on the CTA-pair tcgen05 kernel — ``b200trl_tc_gemm``.

    ``a`` is stored ``[M, K]`` (``a_layout=0``) or ``[K, M]`` (``a_layout=1``), ``b`` ``[N, K]`` or ``[K, N]``
    likewise; bf16, last dim contiguous.  ``accumulate=False``: bf16 ``[M, N]`` result (+ ``bias``);
    ``accumulate=True``: ``out`` (fp32 ``[M, N]``) ``+= D``; ``out_fp32=True``: fp32 result, plain store;
    ``addend``: fp32 ``[M, N]`` added before the bf16 rounding.  ``split_k``: hand the kernel an fp32 scratch so that
    it may split K when the output has too few tiles to fill the machine (bf16 output only).
    """
    _need_cuda(a, "a")
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16 or a.dim() != 2 or b.dim() != 2:
        raise TypeError("tc_gemm needs 2-D bf16 operands")
    if a.stride(1) != 1 or b.stride(1) != 1:
        raise ValueError("tc_gemm operands must be contiguous in their last dimension")
    M, K = (a.shape[1], a.shape[0]) if a_layout else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_layout else (b.shape[0], b.shape[1])
    if K != Kb:
        raise ValueError(f"tc_gemm: contraction lengths differ ({K} vs {Kb})")
    if accumulate or out_fp32:
        if out is None and out_fp32 and not accumulate:
            out = torch.empty(M, N, dtype=torch.float32, device=a.device)
        if out is None or out.dtype != torch.float32 or tuple(out.shape) != (M, N) or out.stride(1) != 1:
            raise ValueError("tc_gemm(accumulate=True) needs an fp32 [M, N] `out`")
    elif out is None:
        out = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
    elif out.dtype != torch.bfloat16 or tuple(out.shape) != (M, N) or out.stride(1) != 1:
        raise ValueError("tc_gemm needs a bf16 [M, N] `out`")
    bb = None if bias is None else bias.to(torch.bfloat16).contiguous()
    kind = TC_OUT_F32_ACC if accumulate else (TC_OUT_F32 if out_fp32 else TC_OUT_BF16)
    ad = None
    if addend is not None:
        if addend.dtype != torch.float32 or tuple(addend.shape) != (M, N) or addend.stride(1) != 1:
            raise ValueError("tc_gemm needs an fp32 [M, N] `addend`")
        ad = addend
    ws, ws_bytes = None, 0
    if split_k and kind == TC_OUT_BF16 and ad is None:
        ws_bytes = int(lib.b200trl_tc_gemm_workspace_bytes(M, N, K, kind))
        ws = _workspace(a.device, ws_bytes, "tc_gemm", zero=False) if ws_bytes else None
    check(lib.b200trl_tc_gemm(_ptr(a), int(a_layout), a.stride(0), _ptr(b), int(b_layout), b.stride(0), M, N, K, kind,
                              _ptr(out), out.stride(0), _ptr(bb), _ptr(ad), ad.stride(0) if ad is not None else 0,
                              int(bool(m_fastest)), _ptr(ws), ws_bytes, _stream(a)), "tc_gemm")
    _count(2 if ws is not None else 1)
    return out


def set_seam_gemm_mask(mask: int) -> int:
    """Bit 0 / 1 / 2 = logits / dH / dW GEMM of ``fused_linear_grpo`` on the tcgen05 kernel (clear = cuBLASLt);
    returns the previous mask, ``mask < 0`` only queries."""
    return int(lib.b200trl_set_seam_gemm_mask(int(mask)))


# ------------------------------------------------------------------------------------------------ tracing
# Every operator that reaches the library runs inside an NVTX range ``b200trl.<op>`` (SURVEY §5 tracing / profiling);
# B200TRL_NVTX=0 leaves the functions unwrapped.
from ._nvtx import nvtx_op as _nvtx_op  # noqa: E402

for _name in ("logprob_entropy_fwd", "masked_logprob_fwd", "logprob_bwd", "mask_stats", "grpo_fused_fwd_bwd",
              "grpo_fused_step", "grpo_loss", "entropy_quantile_mask", "group_advantages", "ppo_rewards_gae",
              "ppo_fused_fwd_bwd", "ppo_fused_step", "ppo_loss", "masked_whiten", "rloo_rewards_advantages", "rloo_loss",
              "fused_linear_logprob_fwd", "fused_linear_grpo", "tc_gemm", "rescale_if_needed", "completion_mask",
              "generation_stats",
              "first_true_indices", "truncate_response"):
    if _name in globals():
        globals()[_name] = _nvtx_op(_name)(globals()[_name])
del _name

