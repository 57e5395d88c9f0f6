"""CUDA-graph capture of one autograd-level step of the hot path.

The small configurations of the path are bound by the host, not by the device: at BASELINE configs[0] (B=4, T=256,
V=32 000) the eager GRPO step costs 64 us of which the kernel is 35 us, and PPO's micro-batch step
(ppo_trainer.py:557-605, mb=8, T=512) 257 us eager against 162 us of device work (bench.py `extra.config1` /
`extra.config3_ppo`).  Every C entry of ``include/b200trl.h`` is graph-capturable (no allocation, no synchronisation,
no host read), so the whole forward + backward of such a step can be recorded once and replayed with one launch call.

``GraphedStep`` does that for any callable built from this package's operators:

    x   = torch.empty(B, T, V, dtype=torch.bfloat16, device="cuda", requires_grad=True)
    buf = {"logits": x, "ids": ids, "mask": mask, "adv": adv, "old": old}
    fn  = swh_trl_b200.GRPOLoss(beta=0.0)

    def body(s):                                     # runs ONCE, under capture, on the static buffers
        s["logits"].grad = None
        out = fn(s["logits"], s["ids"], s["mask"], s["adv"], s["old"])
        out.loss.backward()
        return {"loss": out.loss, "metrics": out.metrics, "dlogits": s["logits"].grad}

    step = swh_trl_b200.GraphedStep(body, buf)
    res  = step.replay(logits=new_logits, ids=new_ids)   # copies into the static buffers, one graph launch
    res["dlogits"]                                       # static output tensors, overwritten by the next replay

Shapes, dtypes and every Python-level branch (loss type, schedule, ``grad_scale``) are frozen at capture; a different
shape needs its own ``GraphedStep``.  There is no CPU path: CPU tensors raise.
"""

from typing import Callable, Dict, Mapping

import torch

from . import _nvtx, ops


class GraphedStep:
    """One captured step: ``body(static_inputs) -> {name: tensor}`` recorded in a ``torch.cuda.CUDAGraph``.

    ``static_inputs`` are used in place as the graph's input buffers (they must stay alive and must not be
    reallocated; ``replay(name=tensor)`` copies new values into them on the current stream).  ``body`` is run
    ``warmup`` times eagerly on a side stream first (lazy initialisation — occupancy queries, tensor maps, the caching
    allocator's pools — must not happen under capture) and once under capture; whatever tensors it returns are the
    static outputs.  Gradients that ``body`` leaves in ``.grad`` of a static leaf must be set to ``None`` at the top of
    ``body`` so that the captured backward allocates them from the graph's own pool.
    """

    def __init__(self, body: Callable[[Dict[str, torch.Tensor]], Mapping[str, torch.Tensor]],
                 static_inputs: Dict[str, torch.Tensor], warmup: int = 3):
        if not static_inputs:
            raise ValueError("GraphedStep needs at least one static input tensor")
        for name, t in static_inputs.items():
            if not isinstance(t, torch.Tensor):
                raise TypeError(f"static input {name!r} is not a tensor")
            if not t.is_cuda:
                raise ValueError(f"static input {name!r} is on {t.device}: the B200 path has no CPU fallback")
        self.inputs = static_inputs
        self.device = next(iter(static_inputs.values())).device
        self.graph = torch.cuda.CUDAGraph()
        self.replays = 0
        # the operators' scratch buffers (loss / metric accumulators, GAE exchange rows ...) are baked into the graph by
        # address: this step owns its own set instead of borrowing the per-stream ones (ops.private_workspaces)
        self._workspaces = {}
        with torch.cuda.device(self.device), ops.private_workspaces(self._workspaces):
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(max(int(warmup), 0)):
                    body(self.inputs)
                with torch.cuda.graph(self.graph, stream=side):
                    outputs = body(self.inputs)
            cur.wait_stream(side)
        if not isinstance(outputs, Mapping) or not all(isinstance(v, torch.Tensor) for v in outputs.values()):
            raise TypeError("body must return a mapping of name -> tensor (the step's static outputs)")
        self.outputs = dict(outputs)

    def load(self, **new_inputs: torch.Tensor) -> None:
        """Copy new values into the named static buffers (device or pinned-host sources; asynchronous)."""
        with torch.no_grad():
            for name, src in new_inputs.items():
                if name not in self.inputs:
                    raise KeyError(f"{name!r} is not a static input of this step (have {sorted(self.inputs)})")
                dst = self.inputs[name]
                if tuple(src.shape) != tuple(dst.shape):
                    raise ValueError(f"{name}: shape {tuple(src.shape)} != captured {tuple(dst.shape)} "
                                     "(a graph is specific to its shapes)")
                if src.dtype != dst.dtype:
                    raise ValueError(f"{name}: dtype {src.dtype} != captured {dst.dtype}")
                if src.data_ptr() != dst.data_ptr():
                    dst.copy_(src, non_blocking=True)

    def replay(self, **new_inputs: torch.Tensor) -> Dict[str, torch.Tensor]:
        """``load(**new_inputs)`` then one graph launch on the current stream; returns the static output tensors."""
        if new_inputs:
            self.load(**new_inputs)
        with _nvtx.nvtx_range("b200trl.graphed_step"):
            self.graph.replay()
        self.replays += 1
        return self.outputs

    __call__ = replay
