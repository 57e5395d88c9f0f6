#!/usr/bin/env python
"""BASELINE config 5: GRPO sequence-sharded over N GPUs, B=256 T=4096 V=151936 (G=8), STRONG scaling.

    python tools/bench_config5.py                                        # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_config5.py

The 256 sequences (318.6 GB of bf16 logits) never co-reside: rank r owns global rows [r*256/N, (r+1)*256/N) and
streams them in micro-batches of MB=4 sequences (16 384 logit-tokens = 4.98 GB of logits + 4.98 GB of dlogits,
buffers reused).  Logits of a micro-batch are regenerated on the device OUTSIDE the timed region (a model forward
would produce them); the timed region of a micro-batch is the hot path itself: mask stats -> K1 fused ->
K2 -> autograd hand-back -> packed metric all-gather.  The per-generation-batch step (reward all-gather over NCCL ->
K3 group advantages on the gathered, rank-major order) is timed once and added.  One JSON line on rank 0:
aggregate logit-tokens/s = 256*4096 / max over ranks of (advantage step + sum of micro-batch times).
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

B_GLOBAL, T, V, G, MB = 256, 4096, 151936, 8, 4


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    import swh_trl_b200 as S

    b_local = B_GLOBAL // world
    n_mb = b_local // MB
    passes = int(os.environ.get("C5_PASSES", 1))
    gen = torch.Generator(device=dev)
    rewards_local = torch.empty(b_local, 1, device=dev)
    for i in range(b_local):  # per-sequence seeding: the same global data whatever the rank count
        gen.manual_seed(9_000_003 + rank * b_local + i)
        rewards_local[i] = torch.randn(1, generator=gen, device=dev)
    weights = torch.ones(1, device=dev)
    loss_fn = S.GRPOLoss(beta=0.04, epsilon_low=0.2, epsilon_high=0.2, loss_type="bnpo",
                         importance_sampling_level="token", max_completion_length=T)
    logits = torch.empty(MB, T, V, dtype=torch.bfloat16, device=dev).requires_grad_(True)
    metrics_all = torch.empty(world, 8, device=dev) if world > 1 else None

    def make_microbatch(m):
        """logits / ids / mask / old / ref of micro-batch m of this rank, seeded per global sequence."""
        ids = torch.empty(MB, T, dtype=torch.long, device=dev)
        mask = torch.zeros(MB, T, dtype=torch.int32, device=dev)
        with torch.no_grad():
            for i in range(MB):
                b = rank * b_local + m * MB + i
                gen.manual_seed(1_000_003 + b)
                for t0 in range(0, T, 1024):  # 1024 rows at a time: bounds the fp32 temporary to 0.6 GB
                    logits[i, t0:t0 + 1024] = torch.randn(1024, V, generator=gen, device=dev).to(torch.bfloat16)
                ids[i] = torch.randint(0, V, (T,), generator=gen, device=dev)
                n = int(torch.randint(T // 2, T + 1, (1,), generator=gen, device=dev))
                mask[i, :n] = 1
            lp0, _ = S.logprobs_and_entropy(logits.detach(), ids, 1.0, compute_entropy=False)
            old = lp0 + torch.randn(MB, T, generator=gen, device=dev) * 0.3
            ref = lp0 + torch.randn(MB, T, generator=gen, device=dev) * 0.1
        return ids, mask, old, ref

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # warm-up (NCCL communicator, kernels, allocator)
    ids, mask, old, ref = make_microbatch(0)
    for _ in range(3):
        adv = S.group_advantages(rewards_local, weights, G)["advantages"]
        logits.grad = None
        out = loss_fn(logits, ids, mask, adv[:MB], old, ref)
        out.loss.backward()
        if world > 1:
            dist.all_gather_into_tensor(metrics_all, out.metrics.reshape(1, 8))
    barrier()

    total_ms = 0.0
    for _ in range(passes):
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        adv = S.group_advantages(rewards_local, weights, G)["advantages"]  # NCCL all-gather + K3, once per batch
        e1.record()
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1)
        for m in range(n_mb):
            ids, mask, old, ref = make_microbatch(m)  # untimed: stands in for the model forward
            logits.grad = None
            barrier()  # ranks finish regenerating at different times; the timed region must not wait for that
            e0, e1 = ev(), ev()
            e0.record()
            out = loss_fn(logits, ids, mask, adv[m * MB:(m + 1) * MB], old, ref)
            out.loss.backward()
            if world > 1:
                dist.all_gather_into_tensor(metrics_all, out.metrics.reshape(1, 8))
            e1.record()
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t) / passes
        print(json.dumps({
            "config": "configs[4]: GRPO sequence-sharded, B=256 T=4096 V=151936, G=8, NCCL reward all-gather",
            "n_gpus": world, "scaling": "strong", "micro_batch_sequences": MB, "micro_batches_per_rank": n_mb,
            "ms_hot_path_per_batch": ms, "value": B_GLOBAL * T / (ms * 1e-3), "unit": "logit-tokens/s",
            "frac_of_measured_hbm_peak_per_gpu": (4 * V * B_GLOBAL * T / world) / (ms * 1e-3) / 1e9 / 6546.6,
            "timing": "CUDA events per micro-batch (synchronised between micro-batches; logits regeneration untimed), "
                      "max over ranks", "loss_last": float(out.loss.detach())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
