"""swh-trl_b200 — the TRL per-token policy-loss hot path on B200 (sm_100a).

Host-side mirror of the reference's call surface over ``libb200trl.so`` (C-ABI in ``include/b200trl.h``).
Import as ``swh_trl_b200``.  No CPU fallback, no Triton, no backend dispatch: importing without the built CUDA
library raises.
"""

from . import _lib  # noqa: F401  (raises ImportError if libb200trl.so is missing)
from ._lib import K1_AUTO, K1_RESIDENT, K1_ROW, B200TRLError, set_k1_path, set_skip_masked  # noqa: F401
from .advantages import generation_metrics, group_advantages  # noqa: F401
from .functional import (  # noqa: F401
    entropy_from_logits,
    fused_linear_logprobs,
    get_high_entropy_mask,
    logprobs_and_entropy,
    masked_mean,
    masked_selective_log_softmax,
    masked_var,
    masked_whiten,
    selective_log_softmax,
    sequence_logps,
)
from .graphs import GraphedStep  # noqa: F401
from .grpo import GRPOLoss, GRPOLossOutput, compute_loss, get_per_token_logps_and_entropies  # noqa: F401
from .liger_seam import B200FusedLinearGRPOLoss  # noqa: F401
from .masks import (  # noqa: F401
    completion_mask_from_eos,
    first_true_indices,
    truncate_response,
    truncate_response_with_lengths,
)
from .patch import patch_trl  # noqa: F401
from .ppo import INVALID_LOGPROB, PPOLossOutput, ppo_loss, ppo_rewards_gae  # noqa: F401
from .rloo import RLOOLossOutput, rloo_loss, rloo_rewards_advantages  # noqa: F401

__version__ = "0.1.0"
