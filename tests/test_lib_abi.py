"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol
that include/b200trl.h declares; argument validation that needs no launch returns the documented codes."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "b200trl.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200trl_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from swh_trl_b200 import _lib
    names = _declared()
    assert len(names) >= 19
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/b200trl.h but not exported"
    assert set(names) == set(_lib.PROTOTYPES), set(names) ^ set(_lib.PROTOTYPES)


def test_version_and_error_text():
    from swh_trl_b200 import _lib
    assert _lib.lib.b200trl_version() >= 100
    assert _lib.lib.b200trl_grpo_loss_workspace_bytes(16) >= 16 * 8 * 4
    assert _lib.lib.b200trl_ppo_gae_workspace_bytes(64, 512) > 0


def test_argument_validation_without_gpu():
    """Null pointers / bad enums are rejected before any CUDA call (status -1 / -2 + message)."""
    from swh_trl_b200 import _lib
    lib = _lib.lib
    rc = lib.b200trl_logprob_entropy_fwd(None, 0, 4, 1024, 1024, 0, 0, None, 1.0, None, None, None, None)
    assert rc == -1 and b"null" in lib.b200trl_last_error()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = lib.b200trl_logprob_entropy_fwd(p, 9, 4, 1024, 1024, 0, 0, p, 1.0, p, None, None, None)
    assert rc == -2
    rc = lib.b200trl_logprob_entropy_fwd(p, 0, 4, 1024, 8, 0, 0, p, 1.0, p, None, None, None)  # stride < vocab
    assert rc == -1
    rc = lib.b200trl_logprob_entropy_fwd(p, 0, 4, 1024, 1024, 3, 0, p, 1.0, p, None, None, None)  # 4 % 3 != 0
    assert rc == -1 and b"rows_per_batch" in lib.b200trl_last_error()
    rc = lib.b200trl_group_advantages(p, p, 10, 1, 4, 1, 0, 10, p, p, p, p, p, p, None)  # 10 % 4 != 0
    assert rc == -1 and b"multiple" in lib.b200trl_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "group_advantages")
    with pytest.raises(NotImplementedError):
        _lib.check(-2, "x")
    assert lib.b200trl_set_k1_path(7) == -1
    # the production build carries no trace hooks (they cost the hot kernel 20 %): the entry point says so
    assert lib.b200trl_k1_set_trace(None, 0) == -2 and b"trace" in lib.b200trl_last_error()
    # the seam call validates before it touches cuBLASLt or the device
    assert lib.b200trl_fused_linear_grpo_workspace_bytes(8, 2048, 3584, 152064, 2) > 2 * 2048 * 152064 * 2
    rc = lib.b200trl_fused_linear_grpo(None, None, None, 1, 1, 8, 8, None, None, None, None, None, None, 1.0, 1, None,
                                       None, None, None, None, None, None, None, None, None)
    assert rc == -1 and b"null" in lib.b200trl_last_error()
    rc = lib.b200trl_completion_mask(None, 2, 4, 0, None, None, None)
    assert rc == -1
    # the one-launch step: its outputs and workspace are mandatory, the mask statistics may be left to the call (both
    # pointers null) but not half-given; the plain fused entry point still needs them
    cfg = _lib.GrpoCfg()
    assert lib.b200trl_grpo_fused_step_workspace_bytes(16) >= 16 + 1024 * 32 + 16 + 64
    assert lib.b200trl_grpo_fused_step_workspace_bytes(5000) >= lib.b200trl_grpo_loss_workspace_bytes(5000) + 16 + 20000
    rc = lib.b200trl_grpo_fused_step(p, 0, 2, 4, 32768, 32768, 0, p, p, p, None, None, ctypes.byref(cfg), 1.0, None, None,
                                     p, p, p, None, 32768, 0, None, p, p, None)
    assert rc == -1 and b"null" in lib.b200trl_last_error()
    rc = lib.b200trl_grpo_fused_step(p, 0, 2, 4, 32768, 32768, 0, p, p, p, None, None, ctypes.byref(cfg), 1.0, p, None,
                                     p, p, p, None, 32768, 0, p, p, p, None)
    assert rc == -1 and b"row_count" in lib.b200trl_last_error()
    rc = lib.b200trl_grpo_fused_fwd_bwd(p, 0, 2, 4, 32768, 32768, 0, p, p, p, None, None, ctypes.byref(cfg), 1.0, None,
                                        None, p, p, p, None, 32768, 0, None)
    assert rc == -1 and b"row_count" in lib.b200trl_last_error()
    rc = lib.b200trl_generation_stats(None, 1, 4, p, 1, p, p, p, 1, p, None)
    assert rc == -1 and b"null" in lib.b200trl_last_error()
    rc = lib.b200trl_generation_stats(p, 0, 4, p, 1, p, p, p, 1, p, None)
    assert rc == -1 and b"sizes" in lib.b200trl_last_error()
    assert lib.b200trl_completion_mask(None, 0, 4, 0, None, None, None) == 0  # empty batch: nothing to launch


def test_no_cpu_fallback():
    import torch
    import swh_trl_b200 as S
    with pytest.raises(RuntimeError, match="no CPU path"):
        S.selective_log_softmax(torch.randn(2, 3, 16), torch.zeros(2, 3, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU path"):
        S.GRPOLoss()(torch.randn(1, 2, 16), torch.zeros(1, 2, dtype=torch.long), torch.ones(1, 2), torch.ones(1))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under swh-trl_b200/ may reference it."""
    pkg = os.path.join(ROOT, "swh-trl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("test oracle", ""), os.path.join(dirpath, f)


def test_k1_geometry_policy():
    """The shape the resident kernel takes per vocabulary and mode (host-only logic): streaming modes
    never need a cluster; the fused pass takes the smallest cluster, then the least tail-chunk waste."""
    from swh_trl_b200 import _lib

    def geom(vocab, mode):
        out = (ctypes.c_int32 * 4)()
        rc = _lib.lib.b200trl_k1_geometry(vocab, mode, out)
        return rc, tuple(out)

    FWD, BWD, FUSED = 0, 1, 2
    for v in (32000, 50304, 151936, 262144, 1 << 20):
        assert geom(v, FWD) == (0, (256, 1, 6, 16384))      # two twin CTAs per SM stream whole rows
        want_slots = 4 if v * 2 >= 200000 else (6 if v == 50304 else 5)  # a short ring: see pick_geom
        assert geom(v, BWD) == (0, (768, 1, want_slots, 24576))
    assert geom(151936, FUSED) == (0, (640, 2, 11, 20480))   # config 2: half a row per CTA, 8 chunks of 20 KB
    assert geom(152064, FUSED) == (0, (640, 2, 11, 20480))   # config 4's vocabulary
    assert geom(100352, FUSED) == (0, (640, 1, 11, 20480))   # the only shape that holds 200 KB in one CTA
    assert geom(50304, FUSED) == (0, (640, 1, 11, 20480))    # 4.9 chunks of 20 KB vs 4.09 of 24 KB
    assert geom(49152, FUSED) == (0, (768, 1, 9, 24576))     # exactly four 24 KB chunks
    assert geom(32000, FUSED) == (0, (256, 1, 6, 16384))     # a whole row fits one twin CTA
    assert geom(262144, FUSED)[1][1] == 4 and geom(524288, FUSED)[1][1] == 8
    assert geom(8192, FUSED)[0] == -2                                # tiny rows go to the row kernel
    # vocab % 8 != 0 (skewed rows): same shapes, with 16 bytes of the ring reserved for the aligned span of a slice
    assert geom(32001, FWD) == (0, (256, 1, 6, 16384)) and geom(50257, FUSED) == (0, (640, 1, 11, 20480))
    assert geom(151937, FUSED) == (0, (640, 2, 11, 20480)) and geom(50257, BWD)[1][0] == 768
    assert geom(1 << 23, FUSED)[0] == -2                             # 16 MB rows exceed an 8-CTA cluster
