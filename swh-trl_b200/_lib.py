"""ctypes binding of ``libb200trl.so`` (the C-ABI declared in ``include/b200trl.h``).

There is no CPU fallback and no other backend: if the shared library is
missing, importing this module raises.  Build it with
``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C swh-trl_b200/csrc``.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200TRL_LIB selects another build of the same library (e.g. the trace build, `make -C swh-trl_b200/csrc trace`)
LIB_PATH = os.environ.get("B200TRL_LIB") or os.path.join(_HERE, "lib", "libb200trl.so")


class B200TRLError(RuntimeError):
    """A C-ABI call returned a negative status."""


class GrpoCfg(C.Structure):
    """``b200trl_grpo_cfg`` (include/b200trl.h)."""

    _fields_ = [
        ("beta", C.c_float),
        ("clip_low", C.c_float),
        ("clip_high", C.c_float),
        ("delta", C.c_float),
        ("has_delta", C.c_int32),
        ("loss_type", C.c_int32),
        ("is_level", C.c_int32),
        ("max_completion_length", C.c_float),
        ("grad_scale", C.c_float),
        ("skip_masked", C.c_int32),
    ]


DTYPE_BF16, DTYPE_F16, DTYPE_F32, DTYPE_F64 = 0, 1, 2, 3
LOSS_TYPES = {"grpo": 0, "bnpo": 1, "dr_grpo": 2}
IS_LEVELS = {"token": 0, "sequence": 1}
KL_ESTIMATORS = {"k1": 0, "k3": 1}
K1_AUTO, K1_ROW, K1_RESIDENT = 0, 1, 2
NUM_GRPO_METRICS = 8
NUM_PPO_STATS = 8

_p, _i64, _i32, _f = C.c_void_p, C.c_int64, C.c_int, C.c_float

# name -> (restype, argtypes); every symbol include/b200trl.h declares
PROTOTYPES = {
    "b200trl_version": (C.c_int, []),
    "b200trl_last_error": (C.c_char_p, []),
    "b200trl_set_k1_path": (C.c_int, [_i32]),
    "b200trl_set_skip_masked": (C.c_int, [_i32]),
    "b200trl_k1_geometry": (C.c_int, [_i64, _i32, _p]),
    "b200trl_k1_set_trace": (C.c_int, [_p, _i64]),
    "b200trl_logprob_entropy_fwd": (C.c_int, [_p, _i32, _i64, _i64, _i64, _i64, _i64, _p, _f, _p, _p, _p, _p]),
    "b200trl_masked_logprob_fwd": (C.c_int, [_p, _i32, _i64, _i64, _i64, _i64, _i64, _p, _p, _f, _p, _p, _p, _p]),
    "b200trl_logprob_bwd": (C.c_int, [_p, _i32, _i64, _i64, _i64, _i64, _i64, _p, _f, _p, _p, _p, _i64, _i64, _p]),
    "b200trl_mask_stats": (C.c_int, [_p, _i64, _i64, _p, _p, _p]),
    "b200trl_grpo_fused_fwd_bwd": (C.c_int, [_p, _i32, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p,
                                             C.POINTER(GrpoCfg), _f, _p, _p, _p, _p, _p, _p, _i64, _i64, _p]),
    "b200trl_grpo_fused_step_workspace_bytes": (_i64, [_i64]),
    "b200trl_grpo_fused_step": (C.c_int, [_p, _i32, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p,
                                          C.POINTER(GrpoCfg), _f, _p, _p, _p, _p, _p, _p, _i64, _i64, _p, _p, _p, _p]),
    "b200trl_grpo_loss_workspace_bytes": (_i64, [_i64]),
    "b200trl_grpo_loss": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _i64, C.POINTER(GrpoCfg), _p, _p, _p, _p, _p,
                                    _p, _p]),
    "b200trl_entropy_quantile_workspace_bytes": (_i64, [_i64]),
    "b200trl_entropy_quantile_mask": (C.c_int, [_p, _p, _i64, _f, _p, _p, _p, _p]),
    "b200trl_group_advantages": (C.c_int, [_p, _p, _i64, _i64, _i64, _i32, _i64, _i64, _p, _p, _p, _p, _p, _p, _p]),
    "b200trl_generation_stats": (C.c_int, [_p, _i64, _i64, _p, _i64, _p, _p, _p, _i64, _p, _p]),
    "b200trl_ppo_gae_workspace_bytes": (_i64, [_i64, _i64]),
    "b200trl_ppo_rewards_gae": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i64, _f, _i32, _f, _f, _i32, _p, _p, _p, _p, _p,
                                          _p, _p, _p]),
    "b200trl_ppo_fused_fwd_bwd": (C.c_int, [_p, _i32, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _f, _f, _f, _p,
                                            _p, _p, _p, _i64, _i64, _p]),
    "b200trl_ppo_fused_step_workspace_bytes": (_i64, [_i64]),
    "b200trl_ppo_fused_step": (C.c_int, [_p, _i32, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _f, _f, _f, _f,
                                         _f, _p, _p, _p, _p, _i64, _i64, _p, _p, _p, _p]),
    "b200trl_ppo_loss": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _f, _f, _f, _f, _p, _p, _p, _p]),
    "b200trl_rloo_rewards_advantages": (C.c_int, [_p, _p, _p, _p, _i64, _i64, _f, _i64, _i32, _f, _i32, _i32, _p, _p, _p,
                                                  _p, _p, _p]),
    "b200trl_rloo_loss": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i64, _f, _f, _p, _p, _p, _p]),
    "b200trl_fused_linear_workspace_bytes": (_i64, [_i64, _i64]),
    "b200trl_fused_linear_logprob_fwd": (C.c_int, [_p, _i64, _p, _i64, _i64, _i64, _i64, _p, _f, _p, _p, _p, _p, _p]),
    "b200trl_tc_gemm_workspace_bytes": (_i64, [_i64, _i64, _i64, _i32]),
    "b200trl_tc_gemm": (C.c_int, [_p, _i32, _i64, _p, _i32, _i64, _i64, _i64, _i64, _i32, _p, _i64, _p, _p, _i64, _i32, _p,
                                  _i64, _p]),
    "b200trl_set_seam_gemm_mask": (C.c_int, [_i32]),
    "b200trl_fused_linear_grpo_workspace_bytes": (_i64, [_i64, _i64, _i64, _i64, _i64]),
    "b200trl_fused_linear_grpo": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p, C.POINTER(GrpoCfg),
                                            _f, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "b200trl_fused_linear_grpo_trimmed": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p,
                                                    C.POINTER(GrpoCfg), _f, C.POINTER(C.c_int64), _p, _p, _p, _p, _p, _p,
                                                    _p, _p, _p, _p]),
    "b200trl_masked_workspace_bytes": (_i64, [_i64]),
    "b200trl_masked_whiten": (C.c_int, [_p, _p, _i64, _i32, _p, _p, _p, _p]),
    "b200trl_first_true_indices": (C.c_int, [_p, _i64, _i64, _p, _p]),
    "b200trl_completion_mask": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _p]),
    "b200trl_truncate_response": (C.c_int, [_p, _i64, _i64, _i32, _i64, _i64, _p, _p, _p]),
    "b200trl_rescale_if_needed": (C.c_int, [_p, _i32, _i64, _i64, _i64, _p, _f, _p]),
    "b200trl_rescale_if_needed_batched": (C.c_int, [_p, _i32, _i64, _i64, _i64, _i64, _i64, _p, _f, _p]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the B200 CUDA library has not been built and there is no CPU fallback. "
            "Run `make -C swh-trl_b200/csrc` (needs nvcc with sm_100a support)."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the build is stale
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(status: int, what: str) -> None:
    """Turn a negative C status into the Python exception the reference would raise."""
    if status >= 0:
        return
    msg = (lib.b200trl_last_error() or b"").decode()
    if status == -1:
        raise ValueError(f"{what}: {msg}")
    if status == -2:
        raise NotImplementedError(f"{what}: {msg}")
    raise B200TRLError(f"{what}: {msg}")


def set_skip_masked(on: bool) -> bool:
    """Opt-in: do not read rows the loss ignores in the fused passes (see ``b200trl_set_skip_masked``)."""
    return bool(lib.b200trl_set_skip_masked(int(bool(on))))


def set_k1_path(path: int) -> int:
    """Select the K1 implementation (K1_AUTO / K1_ROW / K1_RESIDENT); returns the previous one."""
    prev = lib.b200trl_set_k1_path(path)
    if prev < 0:
        raise ValueError(f"unknown K1 path {path}")
    return prev
