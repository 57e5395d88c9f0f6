"""NVTX ranges around the hot-path operators (SURVEY.md §5 tracing / profiling): they show up on an nsys / ncu
timeline as ``b200trl.<op>`` and cost nothing when no profiler is attached.  ``B200TRL_NVTX=0`` removes even
the push / pop calls."""

from __future__ import annotations

import contextlib
import os

import torch

_ENABLED = os.environ.get("B200TRL_NVTX", "1") != "0"


@contextlib.contextmanager
def nvtx_range(name: str):
    if _ENABLED and torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield


def nvtx_op(name: str):
    """Decorator form: the wrapped operator runs inside a ``b200trl.<name>`` range."""
    def deco(fn):
        if not _ENABLED:
            return fn
        import functools

        label = "b200trl." + name
        push, pop = torch.cuda.nvtx.range_push, torch.cuda.nvtx.range_pop

        @functools.wraps(fn)
        def wrapped(*args, **kwargs):
            # no availability check per call (it cost 7 us a step): these operators raise on CPU tensors anyway, and
            # nvtx push / pop are harmless no-ops without a profiler
            push(label)
            try:
                return fn(*args, **kwargs)
            finally:
                pop()
        return wrapped
    return deco
