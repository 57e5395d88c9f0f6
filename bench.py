#!/usr/bin/env python
"""bench.py — fwd+bwd logit-tokens/s of the fused logprob + GRPO loss hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One JSON line on stdout (rank 0).  A *step* is one pass of the hot path over one batch of BASELINE config 2
(B=16 sequences, T=1024, V=151936 bf16 logits, 8 completions per prompt) per GPU:

    [once per generation batch -- 16 loss steps with the fork's own script settings (steps_per_generation 8 x
     num_iterations 2, examples/scripts/grpo_train.py:493-496), as the reference does at grpo_trainer.py:1497:
     reward all-gather (N>1) -> K3 group advantages]  -> mask stats -> K1 fused log-prob / entropy / dlogits with the loss value and the metric
    sums folded into the same launch -> autograd hand-back of dlogits -> packed metric rows kept in a device ring (N>1:
    one all-gather per logging interval)

* ``value``   whole-job logit-tokens/s with inputs resident in HBM (CUDA events, barrier both sides, max over ranks);
* ``e2e``     same metric through the public API starting from PINNED HOST logits (H2D inside the timed region,
              loss + metrics + log-probs read back D2H every step);
* ``roofline``  algorithmic bytes (4*V per logit-token: bf16 read once + bf16 dlogits written once) of the K1
              launch / its CUDA-event duration measured inside the timed region, against MEASURED_PEAKS.json;
* ``cpu_baseline``  the CPU oracle (torch restatement of the reference path, fp32) on a bounded sample, rank 0, N=1.

``--impl reference`` times the reference's own CPU implementation alone: ``GRPOTrainer._compute_loss`` as lifted
verbatim into ``oracle/_ref`` by ``oracle/build_ref.py`` (``kind: "reference"``); if that directory did not travel, the
vectorised port ``oracle/trl_oracle.py`` (``kind: "port"``).

The line also carries ``extra`` — every other BASELINE config measured in the same process with CUDA events and the
same clock sampler: config 1, the two-phase (sequence-level IS) step, config 3 (PPO), config 4 (the Liger seam:
ms, TFLOP/s against the measured sustained bf16 peak, and the no-logits tcgen05 forward) and config 5 (B=256, T=4096
strong scaling over the N ranks) — and ``parity_check``: outside the timed region every rank compares its K3
advantages with the CPU oracle's on the gathered rewards, and sampled log-probs with the oracle's.
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

# steps_per_generation / num_iterations: the fork's own training script (examples/scripts/grpo_train.py:493-496:
# gradient_accumulation_steps=8 -> steps_per_generation=8, num_iterations=2): a generation batch feeds 16 loss steps
CFG = dict(B=16, T=1024, V=151936, G=8, beta=0.04, epsilon=0.2, loss_type="bnpo", level="token", temperature=1.0,
           steps_per_generation=8, num_iterations=2)
WORKLOAD = "configs[1]: GRPO loss fwd+bwd, Qwen2.5 vocab V=151936, B=16 T=1024, 8 completions/prompt, per GPU"
METRIC = "fwd+bwd logit-tokens/s for fused logprob+GRPO loss; % of HBM roofline"
UNIT = "logit-tokens/s"


def static_config(n_gpus: int) -> dict:
    """The workload description: identical, key for key, in the B200 arm and the reference arm (run-specific values
    such as the loss or the measured padding fraction live in ``run``)."""
    B, T, V, G = CFG["B"], CFG["T"], CFG["V"], CFG["G"]
    return {"workload": WORKLOAD, "loss_type": CFG["loss_type"], "importance_sampling_level": CFG["level"],
            "beta": CFG["beta"], "epsilon": CFG["epsilon"], "old_per_token_logps": True, "global_batch": B * n_gpus,
            "seq_len": T, "vocab": V, "num_generations": G, "steps_per_generation": CFG["steps_per_generation"],
            "num_iterations": CFG["num_iterations"],
            "parallelism": f"sequence-sharded x{n_gpus}, no V-sized collective",
            "l2": "inputs 4.98 GB per GPU per step >> 126 MB L2, no flush needed"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tflops():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json)"
    return 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """(dram bytes per K1 launch, provenance) from the committed ``ncu --set full`` capture: a STATIC number read from
    ``profiles/k1_traffic.json``, not measured by this run."""
    path = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("dram_bytes_per_launch"), f"static: ncu --set full capture {d.get('captured', 'round 1')} (profiles/k1_traffic.json)"
    return None, "no capture committed"


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

    def summarise(self, t0, t1):
        """Clock / power / throttle summary of the samples taken in [t0, t1] (wall clock)."""
        if self._thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no sampler"]}
        rows = list(self.rows)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["no samples"]}
        inside = [r for r in rows if t0 is not None and t0 <= r[0] <= t1]
        where = "timed region"
        if not inside:  # region shorter than the polling period: take the sample closest to it
            mid = 0.5 * ((t0 or rows[-1][0]) + (t1 or rows[-1][0]))
            inside = [min(rows, key=lambda r: abs(r[0] - mid))]
            where = "nearest sample to the timed region"
        reasons = sorted({n for r in inside for n in r[3]})
        return {"sm_mhz": statistics.median(r[1] for r in inside), "sm_max_mhz": self.max_sm, "reasons": reasons,
                "power_w_max": max(r[2] for r in inside), "samples": len(inside), "window": where, "source": "nvml"}

    def stop(self):
        out = self.summarise(self.t0, self.t1)
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        return out


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_reference_rate(logits_bf16_cpu, ids, mask, adv, old, ref, steps, warmup):
    """``(logit-tokens/s, ms/step, kind)`` of the reference's CPU path fwd+bwd on ``[b, T, V]`` samples, all host threads.

    ``kind == "reference"``: the reference's own ``GRPOTrainer._compute_loss`` (its per-row Python loops and temporaries),
    lifted verbatim into ``oracle/_ref`` by ``oracle/build_ref.py`` and driven by ``oracle/ref_runner.py``.
    ``kind == "port"``: the vectorised restatement ``oracle/trl_oracle.py`` (when ``oracle/_ref`` did not travel).
    Either way the fp32 path: the bf16 logits are upcast outside the timed step (in the reference they come out of the
    model), the step is loss forward + backward to the logits gradient."""
    from oracle import ref_runner as R
    from oracle import trl_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    nb, T = logits_bf16_cpu.shape[0], logits_bf16_cpu.shape[1]
    kind = "reference" if R.available() else "port"
    if kind == "reference":
        mod = R.load()
        trainer = R.make_trainer(mod, beta=CFG["beta"], epsilon_low=CFG["epsilon"], epsilon_high=CFG["epsilon"], delta=None,
                                 loss_type=CFG["loss_type"], importance_sampling_level=CFG["level"],
                                 max_completion_length=T, temperature=CFG["temperature"])
    else:
        cfg = O.GRPOConfigLite(beta=CFG["beta"], epsilon_low=CFG["epsilon"], epsilon_high=CFG["epsilon"],
                               loss_type=CFG["loss_type"], importance_sampling_level=CFG["level"],
                               max_completion_length=T, temperature=CFG["temperature"])
    times = []
    for s in range(warmup + steps):
        b = s % nb
        if kind == "reference":
            x = R.model_logits(logits_bf16_cpu[b:b + 1].float()).requires_grad_(True)
            t0 = time.perf_counter()
            loss = R.compute_loss(trainer, mod, x, ids[b:b + 1], mask[b:b + 1], adv[b:b + 1], old[b:b + 1], ref[b:b + 1])
        else:
            x = logits_bf16_cpu[b:b + 1].float().requires_grad_(True)
            t0 = time.perf_counter()
            loss, _, _, _ = O.grpo_compute_loss(x, ids[b:b + 1], mask[b:b + 1], adv[b:b + 1], cfg, old[b:b + 1],
                                                ref[b:b + 1])
        loss.backward()
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
        del x, loss
    total = sum(times)
    return T * len(times) / total, total / len(times) * 1e3, kind


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores (rank 0 only)."""
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
    rate, ms, kind = cpu_reference_rate(logits, ids, mask, adv, old, ref, args.steps, args.warmup)
    cores = os.cpu_count() or 1
    sample = (f"each step = 1 sequence (T={T}, V={V}) of the config-2 batch, fp32 torch path fwd+bwd, "
              + ("the reference's own GRPOTrainer._compute_loss (oracle/_ref)" if kind == "reference"
                 else "vectorised port (oracle/trl_oracle.py)"))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": static_config(args.gpus),
        "run": {"reference_path": "CPU torch eager, grpo_trainer.py:2058-2137 (" + kind + ")", "host_threads": cores},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ helpers of the GPU arm
def event_ms(fn, iters, warmup=3):
    """Mean CUDA-event milliseconds of ``fn`` over ``iters`` back-to-back calls after ``warmup`` untimed ones."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (t0, time.time())


def graph_replay(S, body, static, iters, flush=None, ms_flush=0.0):
    """``body`` (an autograd-level step: forward + backward) captured once through the package's public
    ``swh_trl_b200.GraphedStep`` (``static`` = the step's input buffers) and replayed ``iters`` times: what the device
    needs when the host is out of the way.  Returns a dict for the bench line; never raises."""
    try:
        step = S.GraphedStep(lambda s: (body(), {})[1], static, warmup=1)

        def replay():
            if flush is not None:
                flush.add_(1.0)
            step.replay()
        ms, _ = event_ms(replay, iters)
        return {"us_per_step": (ms - ms_flush) * 1e3,
                "what": "the same autograd-level step (forward + backward) captured once in a CUDA graph "
                        "(swh_trl_b200.GraphedStep) and replayed"}
    except Exception as e:  # noqa: BLE001 -- an extra must not take the bench line down
        torch.cuda.synchronize()
        return {"error": f"{type(e).__name__}: {e}"[:200]}


def gpu_affinity(local_rank):
    """CPUs and NUMA node next to this rank's GPU (sysfs of its PCI function); (None, None) if the box does not say."""
    try:
        import pynvml as N
        N.nvmlInit()
        try:  # CUDA_VISIBLE_DEVICES may renumber: resolve through the UUID
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
            h = N.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
        except Exception:  # noqa: BLE001
            h = N.nvmlDeviceGetHandleByIndex(local_rank)
        bdf = N.nvmlDeviceGetPciInfo(h).busId
        bdf = bdf.decode() if isinstance(bdf, bytes) else str(bdf)
    except Exception:  # noqa: BLE001
        return None, None
    bdf = bdf.lower()
    if len(bdf.split(":")[0]) == 8:  # NVML pads the domain to 8 hex digits, sysfs uses 4
        bdf = bdf[4:]
    base = f"/sys/bus/pci/devices/{bdf}"
    try:
        node = int(open(base + "/numa_node").read())
        cpus = open(base + "/local_cpulist").read().strip()
    except OSError:
        return None, None
    ids = set()
    for part in cpus.split(","):
        if "-" in part:
            a, b = part.split("-")
            ids.update(range(int(a), int(b) + 1))
        elif part:
            ids.add(int(part))
    return (node if node >= 0 else None), (sorted(ids) or None)


def pinned_near_gpu(shape, dtype, local_rank):
    """A pinned host tensor whose pages are first touched on the CPUs next to this rank's GPU: the thread is bound to
    them while ``cudaHostAlloc`` allocates and pins, so that on a multi-socket host the H2D copies of the N ranks do
    not all cross to one socket's memory.  Returns ``(tensor, {numa_node, cpus})``."""
    node, cpus = gpu_affinity(local_rank)
    info = {"numa_node": node, "cpus_near_gpu": (f"{cpus[0]}-{cpus[-1]}" if cpus else None)}
    old = None
    if cpus:
        try:
            old = os.sched_getaffinity(0)
            os.sched_setaffinity(0, set(cpus) & old or old)
        except OSError:
            old = None
    try:
        t = torch.empty(shape, dtype=dtype, pin_memory=True)
        t.view(torch.uint8).reshape(-1)[::4096].fill_(0)  # touch every page from here
    finally:
        if old is not None:
            os.sched_setaffinity(0, old)
    return t, info


# ------------------------------------------------------------------------------------------------ extras (other configs)
def extra_config1(S, ops, dev, hbm_peak):
    """BASELINE configs[0] on the GPU: B=4, T=256, V=32000, G=4 (the reference's CPU-runnable case)."""
    B, T, V = 4, 256, 32000
    g = torch.Generator(device=dev).manual_seed(11)
    x = torch.randn(B, T, V, generator=g, device=dev).to(torch.bfloat16).requires_grad_(True)
    ids = torch.randint(0, V, (B, T), generator=g, device=dev)
    mask = torch.ones(B, T, dtype=torch.int32, device=dev)
    adv = S.group_advantages(torch.randn(B, 1, generator=g, device=dev), torch.ones(1, device=dev), 4, gathered=True,
                             local_batch=B)["advantages"]
    with torch.no_grad():
        lp0, _ = S.logprobs_and_entropy(x.detach(), ids, 1.0, compute_entropy=False)
    old = lp0 + 0.1
    fn = S.GRPOLoss(beta=0.0, max_completion_length=T)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)  # 256 MB > L2: the 65 MB input would otherwise sit in L2

    def step():
        flush.add_(1.0)
        x.grad = None
        fn(x, ids, mask, adv, old, None).loss.backward()

    def flush_only():
        flush.add_(1.0)
    ms_all, win = event_ms(step, 30)
    ms_flush, _ = event_ms(flush_only, 30)
    ms = ms_all - ms_flush
    out = {"workload": "configs[0]: B=4 T=256 V=32000 G=4, fused logprob + GRPO loss fwd+bwd", "us_per_step": ms * 1e3,
           "logit_tokens_per_s": B * T / (ms * 1e-3), "frac_of_hbm_roofline": 4 * V * B * T / (ms * 1e-3) / 1e9 / hbm_peak,
           "l2": "256 MB flush write between steps (its own time, measured alone, subtracted)"}
    # The eager step above is bound by the host (autograd + ctypes, one launch); the same step captured in a CUDA graph
    # and replayed shows what the device needs.  Same flush between replays.
    def body():
        x.grad = None
        fn(x, ids, mask, adv, old, None).loss.backward()
    out["graph_replay"] = graph_replay(S, body, {"logits": x}, 30, flush, ms_flush)
    if "us_per_step" in out["graph_replay"]:
        out["graph_replay"]["frac_of_hbm_roofline"] = 4 * V * B * T / (out["graph_replay"]["us_per_step"] * 1e-6) / 1e9 / hbm_peak
    return out, win


def extra_config3(S, ops, dev):
    """BASELINE configs[2]: PPO reward shaping + GAE (B=64, T=512) and the clipped policy / value loss."""
    B, T = 64, 512
    g = torch.Generator(device=dev).manual_seed(13)
    lp = -torch.rand(B, T, generator=g, device=dev) * 5
    rlp = -torch.rand(B, T, generator=g, device=dev) * 5
    val = torch.randn(B, T, generator=g, device=dev)
    sc = torch.randn(B, generator=g, device=dev)
    ln = torch.randint(T // 2, T, (B,), generator=g, device=dev)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    out = {}
    n0 = ops.launch_count
    gae = ops.ppo_rewards_gae(lp, rlp, val, sc, ln, 0.05, "k1", 1.0, 0.95, True)
    launches_gae = ops.launch_count - n0

    def gae_step():
        flush.add_(1.0)
        ops.ppo_rewards_gae(lp, rlp, val, sc, ln, 0.05, "k1", 1.0, 0.95, True)

    def flush_only():
        flush.add_(1.0)
    ms_f, _ = event_ms(flush_only, 30)
    ms_g, win = event_ms(gae_step, 30)
    out["rewards_gae_whiten_us"] = (ms_g - ms_f) * 1e3
    out["rewards_gae_launches"] = launches_gae
    newlp = gae["logprobs"] + 0.05 * torch.randn(B, T, generator=g, device=dev)
    ent = torch.rand(B, T, generator=g, device=dev)
    vpred = gae["values"] + 0.1
    n0 = ops.launch_count
    ops.ppo_loss(newlp, gae["logprobs"], gae["advantages"], gae["returns"], gae["values"], vpred, ent, ln, 0.2, 0.2, 0.1)
    launches_loss = ops.launch_count - n0

    def loss_step():
        flush.add_(1.0)
        ops.ppo_loss(newlp, gae["logprobs"], gae["advantages"], gae["returns"], gae["values"], vpred, ent, ln, 0.2, 0.2, 0.1)
    ms_l, _ = event_ms(loss_step, 30)
    out["clipped_losses_us"] = (ms_l - ms_f) * 1e3
    out["clipped_losses_launches"] = launches_loss
    # the V-sized part of a PPO micro-batch (ppo_trainer.py:557-605): mb=8, V=50304 bf16 logits, fwd+bwd
    mb, V = 8, 50304
    x = (torch.randn(mb, T, V, generator=g, device=dev) * 2).to(torch.bfloat16).requires_grad_(True)
    rsp = torch.randint(0, V, (mb, T), generator=g, device=dev)
    vp = (gae["values"][:mb] + 0.1).requires_grad_(True)

    def ppo_step():
        x.grad = None
        vp.grad = None
        S.ppo_loss(x, rsp, gae["logprobs"][:mb], gae["advantages"][:mb], gae["returns"][:mb], gae["values"][:mb], vp,
                   ln[:mb]).loss.backward()
    n0 = ops.launch_count
    ppo_step()
    out["microbatch_step_launches"] = ops.launch_count - n0
    ms_s, _ = event_ms(ppo_step, 100, warmup=10)
    out["microbatch_step_us"] = ms_s * 1e3
    out["microbatch_step_graph_replay"] = graph_replay(S, ppo_step, {"logits": x, "vpred": vp}, 100)
    out["microbatch_step_shape"] = f"mb={mb} T={T} V={V} bf16 (412 MB logits > L2)"
    out["workload"] = "configs[2]: PPO per-token KL reward + GAE (gamma=1, lam=0.95) + whitening, clipped policy/value loss, B=64 T=512"
    out["l2"] = "256 MB flush write before each small-kernel call (its own time subtracted)"
    return out, win


def extra_config4(S, ops, dev):
    """BASELINE configs[3]: chunked lm_head (3584 -> 152064) fused with the GRPO loss, B=8, T=2048, no [B,T,V] logits."""
    B, T, H, V = 8, 2048, 3584, 152064
    g = torch.Generator(device=dev).manual_seed(17)
    hidden = torch.randn(B, T, H, generator=g, device=dev).to(torch.bfloat16)
    W = (torch.randn(V, H, generator=g, device=dev) * 0.02).to(torch.bfloat16)
    ids = torch.randint(0, V, (B, T), generator=g, device=dev)
    lens = torch.randint(T // 2, T + 1, (B,), generator=g, device=dev)
    mask = (torch.arange(T, device=dev).unsqueeze(0) < lens.unsqueeze(1)).int()
    adv = torch.randn(B, generator=g, device=dev)
    burst, sustained, src = measured_tflops()
    # the no-logits forward (K5 on the CTA-pair tcgen05 kernel): old / ref log-prob passes
    out = {"workload": "configs[3]: chunked lm_head (hidden 3584 -> V=152064) fused with online log-softmax on tcgen05, "
                       "B=8 T=2048, no materialised logits"}
    ms_f, win = event_ms(lambda: ops.fused_linear_logprob_fwd(hidden, W, ids, 1.0), 5, warmup=2)
    fl = 2.0 * B * T * H * V
    out["forward_no_logits"] = {"ms": ms_f, "tflops": fl / ms_f / 1e9, "frac_of_bf16_burst_peak": fl / ms_f / 1e9 / burst,
                                "frac_of_bf16_sustained_peak": fl / ms_f / 1e9 / sustained,
                                "kernel": "tc_gemm_kernel<K-major, K-major, statistics> (tcgen05.mma.cta_group::2)"}
    with torch.no_grad():
        lp0, _, _ = ops.fused_linear_logprob_fwd(hidden, W, ids, 1.0, want_entropy=False)
    old = lp0 + torch.randn(B, T, generator=g, device=dev) * 0.3
    ref = lp0 + torch.randn(B, T, generator=g, device=dev) * 0.1
    h = hidden.clone().requires_grad_(True)
    w = W.clone().requires_grad_(True)
    # the config-4 figure does the DENSE work (every row of the padded batch goes through the three GEMMs, as in the
    # reference's Liger operator); the operator's default -- padding trimmed -- is reported separately below
    fn = S.B200FusedLinearGRPOLoss(beta=0.04, loss_type="bnpo", max_completion_length=T, chunk_size=2, trim_padding=False)

    def seam():
        h.grad = None
        w.grad = None
        loss, _ = fn(h, w, ids, mask, adv, None, old, ref)
        loss.backward()
    n0 = ops.launch_count
    seam()
    launches = ops.launch_count - n0
    ms, win2 = event_ms(seam, 5, warmup=1)
    fl3 = 6.0 * B * T * H * V
    names = {0: "cuBLASLt", 1: "tcgen05 K7 (tc_gemm_kernel, cta_group::2)"}
    m = ops.set_seam_gemm_mask(-1)
    out["fwd_bwd"] = {"ms": ms, "tokens_per_s": B * T / (ms * 1e-3), "tflops": fl3 / ms / 1e9,
                      "frac_of_bf16_sustained_peak": fl3 / ms / 1e9 / sustained, "peak_source": src,
                      "gemms": {"logits": names[m & 1], "dH": names[(m >> 1) & 1], "dW": names[(m >> 2) & 1]},
                      "our_launches": launches, "chunk_sequences": 2, "padding": "computed (trim_padding=False)"}
    fn_t = S.B200FusedLinearGRPOLoss(beta=0.04, loss_type="bnpo", max_completion_length=T)  # default: trim_padding=True

    def seam_trimmed():
        h.grad = None
        w.grad = None
        loss, _ = fn_t(h, w, ids, mask, adv, None, old, ref)
        loss.backward()
    ms_t, win3 = event_ms(seam_trimmed, 5, warmup=1)
    kept = float(lens.sum()) / (B * T)
    out["fwd_bwd_padding_trimmed"] = {
        "ms": ms_t, "tokens_per_s_of_the_padded_batch": B * T / (ms_t * 1e-3), "unmasked_row_fraction": kept,
        "tflops_on_the_rows_computed": fl3 * kept / ms_t / 1e9,
        "what": "NOT the config-4 figure: the operator's default leaves the rows behind each sequence's last unmasked "
                "token out of the three contractions (one chunk per sequence, one D2H read of B lengths per call); "
                "same loss / metrics / gradients"}
    win2 = (win2[0], win3[1])
    del h, w, hidden, W
    torch.cuda.empty_cache()
    return out, (win[0], win2[1])


def extra_two_phase(S, x, ids, mask, adv, old, ref, T, V, hbm_peak):
    """The fork's production setting (examples/scripts/grpo_train.py:491-514): sequence-level importance sampling with
    old log-probs -> K1 forward, K2 with the per-token gradient, K1 backward (2R + 1W of the logits)."""
    fn = S.GRPOLoss(beta=CFG["beta"], epsilon_low=CFG["epsilon"], epsilon_high=CFG["epsilon"], loss_type=CFG["loss_type"],
                    importance_sampling_level="sequence", max_completion_length=T, temperature=CFG["temperature"])

    def step():
        x.grad = None
        fn(x, ids, mask, adv, old, ref).loss.backward()
    ms, win = event_ms(step, 10)
    n = ids.numel()
    return {"workload": "config-2 tensors, importance_sampling_level='sequence' with old_per_token_logps (two-phase schedule)",
            "ms_per_step": ms, "logit_tokens_per_s": n / (ms * 1e-3),
            "frac_of_4V_roofline": 4 * V * n / (ms * 1e-3) / 1e9 / hbm_peak,
            "ceiling": "0.667 of the 4V roofline: the logits are read twice"}, win


def extra_dropin_skip(S, x, ids, mask, adv, old, ref, T, V, hbm_peak):
    """NOT the headline: the same config-2 step the way the trainer drop-in (``compute_loss``) runs it -- with
    ``skip_masked_rows=True``, because its per-token tensors never leave the call, the rows with completion_mask == 0 are
    not read (loss, metrics and gradients bit-identical, tests/test_gpu_parity.py::test_per_call_masked_row_skipping).
    The headline ``value`` reads and writes every row, as the reference does."""
    fn = S.GRPOLoss(beta=CFG["beta"], epsilon_low=CFG["epsilon"], epsilon_high=CFG["epsilon"], loss_type=CFG["loss_type"],
                    importance_sampling_level=CFG["level"], max_completion_length=T, temperature=CFG["temperature"],
                    skip_masked_rows=True)

    def step():
        x.grad = None
        fn(x, ids, mask, adv, old, ref).loss.backward()
    ms, win = event_ms(step, 20)
    n = ids.numel()
    masked = 1.0 - float(mask.float().mean())
    return {"workload": "config-2 tensors through GRPOLoss(skip_masked_rows=True), the trainer drop-in's setting; NOT the "
                        "headline (masked rows are not read; they are still written as zeros)",
            "masked_token_fraction": masked, "ms_per_step": ms, "logit_tokens_per_s": n / (ms * 1e-3),
            "bytes_moved_per_step": int((4 - 2 * masked) * V * n)}, win


def extra_config5(S, ops, dist, world, rank, dev, x, ids, mask, old, ref, hbm_peak):
    """BASELINE configs[4]: B=256, T=4096, V=151936, G=8 over the N ranks -- STRONG scaling.  Rank r owns 256/N
    sequences and streams them as micro-batches of 4 sequences (16 384 logit-tokens; the config-2 buffer viewed as
    [4, 4096, V] is reused for every micro-batch: synthetic data, so nothing is regenerated inside the region).  One
    reward all-gather + K3 per batch, no host sync between micro-batches, the packed metric rows of ALL micro-batches
    leave in one exchange at the end."""
    from swh_trl_b200 import distributed as D
    Bg, T5, V, G, MB = 256, 4096, CFG["V"], 8, 4
    b_local = Bg // world
    n_mb = b_local // MB
    gen = torch.Generator(device=dev).manual_seed(5000 + rank)
    rewards_local = torch.randn(b_local, 1, generator=gen, device=dev)
    weights = torch.ones(1, device=dev)
    x5 = x.detach().view(MB, T5, V).requires_grad_(True)
    ids5, mask5, old5, ref5 = ids.view(MB, T5), mask.view(MB, T5), old.view(MB, T5), ref.view(MB, T5)
    fn = S.GRPOLoss(beta=CFG["beta"], epsilon_low=CFG["epsilon"], epsilon_high=CFG["epsilon"], loss_type=CFG["loss_type"],
                    importance_sampling_level="token", max_completion_length=T5, temperature=CFG["temperature"])
    ring = torch.empty(n_mb, 8, device=dev)

    def batch():
        adv = S.group_advantages(rewards_local, weights, G)["advantages"]  # NCCL all-gather + K3, once per batch
        for m in range(n_mb):
            x5.grad = None
            # what HF Trainer does with gradient accumulation: loss / GA, then backward -> the upstream gradient the
            # fused pass has already folded in (grad_scale); nothing is rescaled afterwards
            o = fn(x5, ids5, mask5, adv[m * MB:(m + 1) * MB], old5, ref5, grad_scale=1.0 / n_mb)
            (o.loss / n_mb).backward()
            ring[m].copy_(o.metrics)
        return D.gather_metric_rows(ring) if world > 1 else ring  # ONE exchange for the whole batch

    batch()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    batch()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    return {"workload": "configs[4]: GRPO sequence-sharded, B=256 T=4096 V=151936 G=8, NCCL reward all-gather, strong scaling",
            "n_gpus": world, "scaling": "strong", "sequences_per_rank": b_local, "micro_batches_per_rank": n_mb,
            "micro_batch": "4 sequences x 4096 tokens (16 384 logit-tokens, 4.98 GB logits + 4.98 GB dlogits)",
            "ms_per_batch": ms, "value": Bg * T5 / (ms * 1e-3), "unit": UNIT,
            "frac_of_hbm_roofline_per_gpu": 4 * V * b_local * T5 / (ms * 1e-3) / 1e9 / hbm_peak,
            "exchanges_per_batch": "1 reward all-gather + 1 packed [micro_batches, 8] metric all-gather; no host sync in between",
            "timing": "CUDA events around the whole batch, max over ranks"}, (t0, time.time())


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
    json_fd = None
    if world > 1:
        # stdout carries exactly one JSON line: NCCL writes its version banner (and, with NCCL_DEBUG set, its log) to
        # file descriptor 1 from C, so everything written to fd 1 during the run is sent to stderr and the line goes
        # out through a duplicate of the original descriptor
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    import swh_trl_b200 as S
    from swh_trl_b200 import ops

    B, T, V, G = CFG["B"], CFG["T"], CFG["V"], CFG["G"]
    if args.skip_masked:
        S.set_skip_masked(True)
    logits, ids, mask = synth_device(rank, B, T, V, dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    # Rewards and advantages belong to a GENERATION batch, as in the reference: `_generate_and_score_completions` gathers
    # the rewards of steps_per_generation loss steps at once (grpo_trainer.py:1497), normalises them per group (:1917-1938),
    # each loss step takes its slice, and the buffered batch is walked num_iterations times before the next generation
    # (`generate_every = steps_per_generation * num_iterations`, :1430-1431).  So the reward all-gather + K3 run once per
    # SPG * NIT loss steps, over SPG * B sequences per rank.
    SPG, NIT = CFG["steps_per_generation"], CFG["num_iterations"]
    rewards_local = torch.randn(SPG * B, 1, generator=gen, device=dev)
    weights = torch.ones(1, device=dev)
    adv_gen, step_no = [None], [0]
    with torch.no_grad():
        lp0, _ = S.logprobs_and_entropy(logits, ids, CFG["temperature"], compute_entropy=False)
    old = lp0 + torch.randn(B, T, generator=gen, device=dev) * 0.3
    ref = lp0 + torch.randn(B, T, generator=gen, device=dev) * 0.1
    loss_fn = S.GRPOLoss(beta=CFG["beta"], epsilon_low=CFG["epsilon"], epsilon_high=CFG["epsilon"],
                         loss_type=CFG["loss_type"], importance_sampling_level=CFG["level"], max_completion_length=T,
                         temperature=CFG["temperature"])
    x = logits.requires_grad_(True)
    # N > 1: the packed metric rows wait in a device ring and leave in ONE exchange per logging interval (the product's
    # deferred path, grpo.flush_metrics) instead of one all-gather per step
    ring = torch.empty(max(args.steps, 64), 8, device=dev) if world > 1 else None
    ring_pos = [0]
    k1_events = []

    def flush_ring():
        from swh_trl_b200 import distributed as D
        if world > 1 and ring_pos[0] > 0:
            D.gather_metric_rows_device(ring[:ring_pos[0]])
            ring_pos[0] = 0

    def step(record=False):
        x.grad = None
        j = step_no[0] % (SPG * NIT)
        step_no[0] += 1
        if j == 0:  # a new generation batch: all-gather (N>1) + K3 over its SPG * B sequences per rank
            adv_gen[0] = S.group_advantages(rewards_local, weights, G)["advantages"]
        j %= SPG    # the second iteration walks the same buffered slices again
        adv = adv_gen[0][j * B:(j + 1) * B]
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            orig = ops.grpo_fused_step  # the K1 launch (log-probs, entropies, dlogits, loss + metrics sums)

            def timed(*a, **k):
                e0.record()
                r = orig(*a, **k)
                e1.record()
                return r
            ops.grpo_fused_step = timed
        try:
            out = loss_fn(x, ids, mask, adv, old, ref)
        finally:
            if record:
                ops.grpo_fused_step = orig
                k1_events.append((e0, e1))
        out.loss.backward()
        if world > 1:
            if ring_pos[0] == ring.shape[0]:
                flush_ring()
            ring[ring_pos[0]].copy_(out.metrics)
            ring_pos[0] += 1
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        out = step()
    barrier()
    # pre-burn: >= 1 s of the same step so that the timed region reads steady-state clocks and power even when the
    # driver asks for only 20 steps (33 ms)
    burn_steps, t_burn = 0, time.time()
    while time.time() - t_burn < args.burn_s:
        for _ in range(25):
            step()
        burn_steps += 25
        flush_ring()
        torch.cuda.synchronize()
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
    flush_ring()  # inside the timed region: the metrics of these steps have left the rank when the clock stops
    t_end.record()
    barrier()
    if sampler:
        sampler.mark_end()
    if args.profiler_range:
        torch.cuda.profiler.stop()
    elapsed_ms = t_start.elapsed_time(t_end)
    launches = ops.launch_count - launches0
    k1_ms = statistics.mean(a.elapsed_time(b) for a, b in k1_events)

    # ---- e2e: pinned host logits -> H2D -> loss fwd+bwd -> D2H of loss / metrics / log-probs, every step
    e2e_steps = 0 if args.no_e2e else max(3, min(args.steps, 10))
    e2e_s, copy_s, pin_info = float("nan"), float("nan"), {}
    h2d = B * T * V * 2
    d2h = (8 + 1 + B * T) * 4
    if e2e_steps:
        host_logits, pin_info = pinned_near_gpu((B, T, V), torch.bfloat16, local_rank)
        host_logits.copy_(logits.detach())
        host_small = torch.empty(8 + 1 + B * T, dtype=torch.float32, pin_memory=True)
        dev_logits = torch.empty_like(logits).requires_grad_(True)

        def e2e_step():
            dev_logits.grad = None
            with torch.no_grad():
                dev_logits.copy_(host_logits, non_blocking=True)
            adv = S.group_advantages(rewards_local[:B], weights, G)["advantages"]
            o = loss_fn(dev_logits, ids, mask, adv, old, ref)
            o.loss.backward()
            packed = torch.cat([o.metrics, o.loss.detach().reshape(1), o.per_token_logps.reshape(-1)])
            host_small.copy_(packed, non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the caller needs the loss on the host

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = time.perf_counter() - t0
        # the platform's ceiling for this leg: the same host->device copy alone, all ranks at the same time
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            with torch.no_grad():
                dev_logits.copy_(host_logits, non_blocking=True)
        barrier()
        copy_s = (time.perf_counter() - t0) / 3
        del host_logits, dev_logits

    # ---- parity check, outside every timed region: this rank's K3 advantages against the CPU oracle on the gathered
    # rewards (bound of tests/test_gpu_parity.py::test_group_advantages_golden), and sampled log-probs of the fused pass
    from oracle import trl_oracle as O
    from swh_trl_b200 import distributed as D
    res = S.group_advantages(rewards_local, weights, G)
    full_rewards = D.gather_rewards(rewards_local).cpu()
    nloc = SPG * B  # this rank's sequences of one generation batch
    want_loc, want_all, _, want_std, _, _ = O.group_advantages(full_rewards, torch.ones(1), G, True, rank * nloc, nloc)
    bound = 2e-5 * want_all.abs() + 3e-7 / (want_std.repeat_interleave(G) + 1e-4)
    adv_ok = bool(((res["all"].cpu() - want_all).abs() <= bound).all()) and \
        torch.equal(res["advantages"].cpu(), res["all"].cpu()[rank * nloc:(rank + 1) * nloc])
    rows = slice(0, 64)
    lp_dev = out.per_token_logps[3, rows].cpu()
    lp_cpu = O.selective_log_softmax(logits.detach()[3:4, rows].float().cpu(), ids[3:4, rows].cpu())[0]
    lp_err = float((lp_dev - lp_cpu).abs().max())
    ok = torch.tensor([1.0 if (adv_ok and lp_err <= 1e-5) else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    parity = {"ok": bool(ok.item() == 1.0), "advantages_vs_oracle_all_ranks": True if ok.item() == 1.0 else adv_ok,
              "max_abs_logp_err_sampled_rows": lp_err, "rows_sampled": 64,
              "what": "K3 group advantages of every rank vs the CPU oracle on the NCCL-gathered rewards (values within the "
                      "test bound, local slice bit-equal to the global order); 64 fused-pass log-probs vs the oracle, 1e-5"}

    # ---- the other BASELINE configs, same process, CUDA events, same clock sampler
    hbm_peak, peak_src = measured_peaks()
    extra, windows = {}, {}
    if not args.no_extras:
        adv_now = res["advantages"][:B]
        extra["two_phase_sequence_is"], windows["two_phase_sequence_is"] = extra_two_phase(S, x, ids, mask, adv_now, old,
                                                                                        ref, T, V, hbm_peak)
        extra["trainer_dropin_masked_rows_skipped"], windows["trainer_dropin_masked_rows_skipped"] = extra_dropin_skip(
            S, x, ids, mask, adv_now, old, ref, T, V, hbm_peak)
        extra["config5_strong_scaling"], windows["config5_strong_scaling"] = extra_config5(
            S, ops, dist, world, rank, dev, x, ids, mask, old, ref, hbm_peak)
        if rank == 0:
            extra["config1"], windows["config1"] = extra_config1(S, ops, dev, hbm_peak)
            extra["config3_ppo"], windows["config3_ppo"] = extra_config3(S, ops, dev)
            extra["config4_liger_seam"], windows["config4_liger_seam"] = extra_config4(S, ops, dev)
        if sampler:
            for k, (a, b) in windows.items():
                if k in extra:
                    c = sampler.summarise(a, b)
                    extra[k]["clocks"] = {"sm_mhz": c.get("sm_mhz"), "power_w_max": c.get("power_w_max"),
                                          "reasons": c.get("reasons")}
    clocks = sampler.stop() if sampler else None

    # ---- max over ranks
    t = torch.tensor([elapsed_ms, k1_ms, e2e_s, copy_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, k1_ms, e2e_s, copy_s = [float(v) for v in t]

    if rank == 0:
        tokens_per_step = B * T * world
        value = tokens_per_step * args.steps / (elapsed_ms * 1e-3)
        alg_bytes = 4 * V * B * T  # per K1 launch (one GPU): bf16 logits read once + bf16 dlogits written once
        achieved = alg_bytes / (k1_ms * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": static_config(world),
            "run": {"loss": loss_val, "masked_token_fraction": float(1.0 - mask.float().mean()),
                    "skip_masked_rows": bool(args.skip_masked), "pre_burn_steps": burn_steps, "pre_burn_s": args.burn_s,
                    "metric_readback": "packed metric rows kept in a device ring, one all-gather per logging interval "
                                       "(inside the timed region); host read deferred"},
            "clocks": clocks,
            "e2e": ({"value": tokens_per_step * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                     "h2d_gbs_per_gpu_in_step": h2d / (e2e_s / e2e_steps) / 1e9,
                     "h2d_copy_alone_gbs_per_gpu": h2d / copy_s / 1e9,
                     "platform_ceiling": "h2d_copy_alone = the same pinned-host -> device copy with nothing else in the "
                                         "step, all ranks copying at once (max over ranks): what PCIe / host DRAM allow",
                     "pinned_buffer": pin_info} if e2e_steps else None),
            "gpu_launches": launches,
            "parity_check": parity,
            "roofline": {"bound": "hbm", "kernel": "k1_resident_kernel<fwd,bwd> (fused log-prob + entropy + dlogits)",
                         "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k1_ms,
                         "frac_of_nominal_8TBps": achieved / 8000.0,
                         "kernel_share_of_step": k1_ms / (elapsed_ms / args.steps)},
            "extra": extra,
        }
        if world == 1 and not args.no_cpu_baseline:
            nb, cpu_steps = 4, 30  # ~11 s of CPU work on the box's cores
            cpu_logits = logits.detach()[:nb].cpu()
            adv = res["advantages"]
            rate, ms, kind = cpu_reference_rate(cpu_logits, ids[:nb].cpu(), mask[:nb].cpu(), adv[:nb].cpu(), old[:nb].cpu(),
                                                ref[:nb].cpu(), steps=cpu_steps, warmup=2)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind,
                                    "sample": f"{cpu_steps} timed steps of 1 sequence each (T={T}, V={V}; {nb} distinct "
                                              f"sequences of the same batch in turn), fp32 torch path fwd+bwd, "
                                              f"{ms:.0f} ms/step"}
        if json_fd is not None:
            os.write(json_fd, (json.dumps(line) + "\n").encode())
        else:
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
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs (profiling runs)")
    ap.add_argument("--burn-s", type=float, default=1.0, help="seconds of untimed steps before the timed region")
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
