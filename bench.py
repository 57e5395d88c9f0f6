#!/usr/bin/env python
"""bench.py — fwd+bwd logit-tokens/s of the fused logprob + GRPO loss hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One JSON line on stdout (rank 0).  A *step* is one pass of the hot path over one batch of BASELINE config 2
(B=16 sequences, T=1024, V=151936 bf16 logits, 8 completions per prompt) per GPU:

    reward all-gather (N>1) -> K3 group advantages -> mask stats -> K1 fused log-prob/entropy/dlogits
    -> K2 loss + metrics -> autograd hand-back of dlogits -> packed metric all-gather (N>1)

* ``value``   whole-job logit-tokens/s with inputs resident in HBM (CUDA events, barrier both sides, max over ranks);
* ``e2e``     same metric through the public API starting from PINNED HOST logits (H2D inside the timed region,
              loss + metrics + log-probs read back D2H every step);
* ``roofline``  algorithmic bytes (4*V per logit-token: bf16 read once + bf16 dlogits written once) of the K1
              launch / its CUDA-event duration measured inside the timed region, against MEASURED_PEAKS.json;
* ``cpu_baseline``  the CPU oracle (torch restatement of the reference path, fp32) on a bounded sample, rank 0, N=1.

``--impl reference`` times that CPU path alone (the reference is pure Python and cannot travel to the GPU box;
the oracle port, pinned on the reference's own outputs, is what runs).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(B=16, T=1024, V=151936, G=8, beta=0.04, epsilon=0.2, loss_type="bnpo", level="token", temperature=1.0)
WORKLOAD = "configs[1]: GRPO loss fwd+bwd, Qwen2.5 vocab V=151936, B=16 T=1024, 8 completions/prompt, per GPU"
METRIC = "fwd+bwd logit-tokens/s for fused logprob+GRPO loss; % of HBM roofline"
UNIT = "logit-tokens/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per K1 launch from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("dram_bytes_per_launch")
    return None


# ------------------------------------------------------------------------------------------------ synthetic data
def synth_device(rank, B, T, V, dev, seed=0):
    """Per-sequence seeded bf16 logits (sharding-invariant: sequence b of the global batch is always the same)."""
    logits = torch.empty(B, T, V, dtype=torch.bfloat16, device=dev)
    ids = torch.empty(B, T, dtype=torch.long, device=dev)
    mask = torch.zeros(B, T, dtype=torch.int32, device=dev)
    g = torch.Generator(device=dev)
    for i in range(B):
        b = rank * B + i
        g.manual_seed(seed * 1_000_003 + b)
        x = torch.randn(T, V, generator=g, device=dev)
        idx = torch.randint(0, V, (T,), generator=g, device=dev)
        bump = torch.rand(T, generator=g, device=dev) < 0.5  # peaked rows: +8 at the chosen id
        x[torch.arange(T, device=dev)[bump], idx[bump]] += 8.0
        n = int(torch.randint(T // 2, T + 1, (1,), generator=g, device=dev))
        if b % 16 == 1:
            n = 0
        if b % 16 == 2:
            n = T
        logits[i], ids[i] = x.to(torch.bfloat16), idx
        mask[i, :n] = 1
        del x
    return logits, ids, mask


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).

    A thread polls NVML every 2 ms (in process: nvidia-smi needs seconds to come up on an 8-GPU box and its -lms
    period is coarser than a short timed region); samples carry a timestamp and only those between ``mark_start`` and
    ``mark_end`` are summarised.  Falls back to one ``nvidia-smi`` query per mark if NVML is not importable."""

    REASONS = (("hw_slowdown", "HwSlowdown"), ("hw_thermal_slowdown", "HwThermalSlowdown"),
               ("sw_thermal_slowdown", "SwThermalSlowdown"), ("sw_power_cap", "SwPowerCap"),
               ("hw_power_brake_slowdown", "HwPowerBrakeSlowdown"))

    def __init__(self, cuda_index):
        import threading
        self.rows, self.t0, self.t1, self.err = [], None, None, None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml as N
            N.nvmlInit()
            try:  # CUDA_VISIBLE_DEVICES may renumber: resolve through the UUID
                uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
                self.h = N.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
            except Exception:
                self.h = N.nvmlDeviceGetHandleByIndex(cuda_index)
            self.N = N
            self.max_sm = float(N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
        except Exception as e:  # noqa: BLE001 - any NVML failure just disables the sampler
            self.err = f"nvml unavailable: {type(e).__name__}"

    def _poll(self):
        N = self.N
        while not self._stop.is_set():
            try:
                sm = float(N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM))
                pw = N.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                bits = int(N.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                active = [name for name, sym in self.REASONS if bits & getattr(N, "nvmlClocksEventReason" + sym)]
                self.rows.append((time.time(), sm, pw, active))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.002)

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self._thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no sampler"]}
        self._stop.set()
        self._thread.join(timeout=2)
        rows = list(self.rows)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["no samples"]}
        inside = [r for r in rows if self.t0 is not None and self.t0 <= r[0] <= self.t1]
        where = "timed region"
        if not inside:  # region shorter than the polling period: take the sample closest to it
            mid = 0.5 * ((self.t0 or rows[-1][0]) + (self.t1 or rows[-1][0]))
            inside = [min(rows, key=lambda r: abs(r[0] - mid))]
            where = "nearest sample to the timed region"
        reasons = sorted({n for r in inside for n in r[3]})
        return {"sm_mhz": statistics.median(r[1] for r in inside), "sm_max_mhz": self.max_sm, "reasons": reasons,
                "power_w_max": max(r[2] for r in inside), "samples": len(inside), "window": where, "source": "nvml"}


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_reference_rate(logits_bf16_cpu, ids, mask, adv, old, ref, steps, warmup):
    """logit-tokens/s of the oracle (reference fp32 torch path, all host threads) fwd+bwd on [b,T,V] samples."""
    from oracle import trl_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.GRPOConfigLite(beta=CFG["beta"], epsilon_low=CFG["epsilon"], epsilon_high=CFG["epsilon"],
                           loss_type=CFG["loss_type"], importance_sampling_level=CFG["level"],
                           max_completion_length=CFG["T"], temperature=CFG["temperature"])
    nb = logits_bf16_cpu.shape[0]
    times = []
    for s in range(warmup + steps):
        b = s % nb
        t0 = time.perf_counter()
        x = logits_bf16_cpu[b:b + 1].float().requires_grad_(True)  # the reference fp32 path: upcast, then torch eager
        loss, _, _, _ = O.grpo_compute_loss(x, ids[b:b + 1], mask[b:b + 1], adv[b:b + 1], cfg, old[b:b + 1],
                                            ref[b:b + 1])
        loss.backward()
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
        del x, loss
    tokens = logits_bf16_cpu.shape[1]
    total = sum(times)
    return tokens * len(times) / total, total / len(times) * 1e3


def run_reference(args):
    """--impl reference: the reference's CPU implementation (oracle port) on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, T, V = 2, CFG["T"], CFG["V"]
    gen = torch.Generator().manual_seed(0)
    logits = torch.empty(B, T, V, dtype=torch.bfloat16)
    for b in range(B):
        logits[b] = torch.randn(T, V, generator=gen).to(torch.bfloat16)
    ids = torch.randint(0, V, (B, T), generator=gen)
    mask = (torch.arange(T).unsqueeze(0) < torch.tensor([[T], [3 * T // 4]])).int()
    adv = torch.tensor([0.8, -1.2])
    from oracle import trl_oracle as O
    with torch.no_grad():
        lp0 = torch.cat([O.selective_log_softmax(logits[b:b + 1].float(), ids[b:b + 1]) for b in range(B)])
    old = lp0 + torch.randn(B, T, generator=gen) * 0.3
    ref = lp0 + torch.randn(B, T, generator=gen) * 0.1
    rate, ms = cpu_reference_rate(logits, ids, mask, adv, old, ref, args.steps, args.warmup)
    cores = os.cpu_count() or 1
    sample = f"each step = 1 sequence (T={T}, V={V}) of the config-2 batch, fp32 torch path fwd+bwd"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_path": "CPU torch eager (oracle port of grpo_trainer.py:2058-2137)",
                   "host_threads": cores},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)

    import swh_trl_b200 as S
    from swh_trl_b200 import ops

    B, T, V, G = CFG["B"], CFG["T"], CFG["V"], CFG["G"]
    if args.skip_masked:
        S.set_skip_masked(True)
    logits, ids, mask = synth_device(rank, B, T, V, dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    rewards_local = torch.randn(B, 1, generator=gen, device=dev)
    weights = torch.ones(1, device=dev)
    with torch.no_grad():
        lp0, _ = S.logprobs_and_entropy(logits, ids, CFG["temperature"], compute_entropy=False)
    old = lp0 + torch.randn(B, T, generator=gen, device=dev) * 0.3
    ref = lp0 + torch.randn(B, T, generator=gen, device=dev) * 0.1
    loss_fn = S.GRPOLoss(beta=CFG["beta"], epsilon_low=CFG["epsilon"], epsilon_high=CFG["epsilon"],
                         loss_type=CFG["loss_type"], importance_sampling_level=CFG["level"], max_completion_length=T,
                         temperature=CFG["temperature"])
    x = logits.requires_grad_(True)
    metrics_all = torch.empty(world, 8, device=dev) if world > 1 else None
    k1_events = []

    def step(record=False):
        x.grad = None
        adv = S.group_advantages(rewards_local, weights, G)["advantages"]  # all-gather (N>1) + K3
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            orig = ops.grpo_fused_fwd_bwd

            def timed(*a, **k):
                e0.record()
                r = orig(*a, **k)
                e1.record()
                return r
            ops.grpo_fused_fwd_bwd = timed
        try:
            out = loss_fn(x, ids, mask, adv, old, ref)
        finally:
            if record:
                ops.grpo_fused_fwd_bwd = orig
                k1_events.append((e0, e1))
        out.loss.backward()
        if world > 1:
            dist.all_gather_into_tensor(metrics_all, out.metrics.reshape(1, 8))
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        out = step()
    barrier()
    loss_val = float(out.loss.detach())

    launches0 = ops.launch_count
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.profiler_range:  # ncu --profile-from-start off: only the timed region is captured
        torch.cuda.profiler.start()
    if sampler:
        sampler.mark_start()
    t_start.record()
    for _ in range(args.steps):
        step(record=True)
    t_end.record()
    barrier()
    if sampler:
        sampler.mark_end()
    if args.profiler_range:
        torch.cuda.profiler.stop()
    elapsed_ms = t_start.elapsed_time(t_end)
    launches = ops.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    k1_ms = statistics.mean(a.elapsed_time(b) for a, b in k1_events)

    # ---- e2e: pinned host logits -> H2D -> loss fwd+bwd -> D2H of loss / metrics / log-probs, every step
    host_logits = torch.empty(B, T, V, dtype=torch.bfloat16, pin_memory=True)
    host_logits.copy_(logits.detach())
    host_small = torch.empty(8 + 1 + B * T, dtype=torch.float32, pin_memory=True)
    dev_logits = torch.empty_like(logits).requires_grad_(True)
    h2d = host_logits.numel() * 2
    d2h = host_small.numel() * 4

    def e2e_step():
        dev_logits.grad = None
        with torch.no_grad():
            dev_logits.copy_(host_logits, non_blocking=True)
        adv = S.group_advantages(rewards_local, weights, G)["advantages"]
        o = loss_fn(dev_logits, ids, mask, adv, old, ref)
        o.loss.backward()
        packed = torch.cat([o.metrics, o.loss.detach().reshape(1), o.per_token_logps.reshape(-1)])
        host_small.copy_(packed, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller needs the loss on the host

    e2e_steps = 0 if args.no_e2e else max(3, min(args.steps, 10))
    e2e_s = float("nan")
    if e2e_steps:
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = time.perf_counter() - t0
    del host_logits, dev_logits

    # ---- max over ranks
    t = torch.tensor([elapsed_ms, k1_ms, e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, k1_ms, e2e_s = [float(v) for v in t]

    if rank == 0:
        tokens_per_step = B * T * world
        value = tokens_per_step * args.steps / (elapsed_ms * 1e-3)
        peak, peak_src = measured_peaks()
        alg_bytes = 4 * V * B * T  # per K1 launch (one GPU): bf16 logits read once + bf16 dlogits written once
        achieved = alg_bytes / (k1_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "loss_type": CFG["loss_type"], "importance_sampling_level": CFG["level"],
                       "beta": CFG["beta"], "epsilon": CFG["epsilon"], "old_per_token_logps": True,
                       "global_batch": B * world, "seq_len": T, "vocab": V, "num_generations": G,
                       "parallelism": f"sequence-sharded x{world}, no V-sized collective",
                       "l2": "inputs 4.98 GB per GPU per step >> 126 MB L2, no flush needed",
                       "metric_readback": "packed metrics all-gathered on device every step; host read deferred",
                       "skip_masked_rows": bool(args.skip_masked),
                       "masked_token_fraction": float(1.0 - mask.float().mean()),
                       "loss": loss_val},
            "clocks": clocks,
            "e2e": ({"value": tokens_per_step * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h, "steps": e2e_steps} if e2e_steps else None),
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "k1_resident_kernel<fwd,bwd> (fused log-prob + entropy + dlogits)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k1_ms,
                         "frac_of_nominal_8TBps": achieved / 8000.0,
                         "kernel_share_of_step": k1_ms / (elapsed_ms / args.steps)},
        }
        if world == 1 and not args.no_cpu_baseline:
            nb, cpu_steps = 4, 30  # ~11 s of CPU work on the box's cores
            cpu_logits = logits.detach()[:nb].cpu()
            adv = S.group_advantages(rewards_local, weights, G)["advantages"]
            rate, ms = cpu_reference_rate(cpu_logits, ids[:nb].cpu(), mask[:nb].cpu(), adv[:nb].cpu(), old[:nb].cpu(),
                                          ref[:nb].cpu(), steps=cpu_steps, warmup=2)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": f"{cpu_steps} timed steps of 1 sequence each (T={T}, V={V}; {nb} distinct "
                                              f"sequences of the same batch in turn), fp32 torch path fwd+bwd, "
                                              f"{ms:.0f} ms/step"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--profiler-range", action="store_true", help="cudaProfilerStart/Stop around the timed region")
    ap.add_argument("--skip-masked", action="store_true",
                    help="opt-in: do not read rows with completion_mask == 0 (reported in config; NOT the default)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
