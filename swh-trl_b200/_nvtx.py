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

        @functools.wraps(fn)
        def wrapped(*args, **kwargs):
            if not torch.cuda.is_available():
                return fn(*args, **kwargs)
            torch.cuda.nvtx.range_push("b200trl." + name)
            try:
                return fn(*args, **kwargs)
            finally:
                torch.cuda.nvtx.range_pop()
        return wrapped
    return deco
