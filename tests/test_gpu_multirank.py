"""On-hardware N>1 parity: NCCL reward all-gather -> K3 group advantages -> process_slice, and the packed metric
exchange, on real GPUs (one process per GPU, spawned here) against the reference's per-rank outputs
(tests/golden/advantages.pt, produced by the reference's own source at grpo_trainer.py:1497, 1917-1938) and the
oracle's metric block (:2139-2172).  Skipped when the box has fewer GPUs than the case needs; the world-2 gloo tests
(tests/test_dist_gloo.py) cover the same host logic on CPU with the oracle standing in for K3."""
import os
import socket

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _adv_bound(c, want):
    ref_all = want["all_process_advantages"]
    std_r = want["std_grouped_rewards"] if c["scale_rewards"] else torch.full_like(ref_all, 1.0 - 1e-4)
    return ref_all, (2e-5 * ref_all.abs() + 3e-7 / (std_r + 1e-4))


def _worker(rank, world, port, case_idx, q):
    import sys
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import swh_trl_b200 as S
        from oracle import trl_oracle as O
        from swh_trl_b200 import distributed as D
        from tests.conftest import load_golden

        c = load_golden("advantages.pt")[case_idx]
        assert c["world"] == world
        n_local = c["B_global"] // world
        local = c["rewards_per_func"][rank * n_local:(rank + 1) * n_local].clone().to(dev)
        # ---- NCCL all-gather (rank-major) + K3 + this rank's slice, exactly the call the trainer makes
        out = S.group_advantages(local, c["weights"].to(dev), c["G"], c["scale_rewards"])
        want = c["per_rank"][rank]
        ref_all, bound = _adv_bound(c, want)
        got_all = out["all"].cpu()
        assert torch.equal(got_all.isnan(), ref_all.isnan())
        assert bool(((got_all - ref_all).abs().nan_to_num(0.0) <= bound.nan_to_num(1.0)).all()), "advantage values"
        off, cnt = D.process_slice(n_local)
        assert (off, cnt) == (rank * n_local, n_local)
        assert torch.equal(out["advantages"].cpu(), got_all[off:off + cnt]), "local slice is the global one"
        assert torch.equal(out["advantages"].cpu().isnan(), want["advantages"].isnan())
        assert bool(((out["advantages"].cpu() - want["advantages"]).abs().nan_to_num(0.0)
                     <= bound[off:off + cnt].nan_to_num(1.0)).all()), "per-rank advantages vs the reference"
        assert torch.equal(out["is_std_zero"].cpu(), want["is_std_zero"])
        gathered = D.gather_rewards(local).cpu()
        assert torch.equal(gathered.nan_to_num(7.0), c["rewards_per_func"].nan_to_num(7.0)), "rank-major gather"

        # ---- the logging block behind the advantages (grpo_trainer.py:1942-1970): one packed NCCL gather + one launch
        gm_cases = [g for g in load_golden("generation_metrics.pt") if g["world"] == world]
        for gc in gm_cases:
            nl = gc["B_global"] // world
            sl = slice(rank * nl, (rank + 1) * nl)
            a2 = S.group_advantages(gc["rewards_per_func"][sl].clone().to(dev), gc["weights"].to(dev), gc["G"], True)
            full = D.gather_rewards(gc["rewards_per_func"][sl].clone().to(dev))
            gm = S.generation_metrics(gc["attention_mask"][sl].long().to(dev), gc["completion_lengths"][sl].to(dev),
                                      gc["terminated"][sl].to(dev), full, a2["mean"], a2["std"], a2["is_std_zero"],
                                      gc["names"])
            assert gm["num_tokens"] == gc["metrics"]["num_tokens"]
            for k, w in gc["metrics"].items():
                if k != "num_tokens":
                    assert (gm[k] != gm[k] and w != w) or gm[k] == pytest.approx(w, rel=2e-6, abs=1e-6), (k, gm[k], w)

        # ---- packed metric exchange: each rank runs the loss on its own shard, one [world, 8] all-gather
        B, T, V, G = 4, 32, 4096, 2
        logits, ids, mask = O.synth_batch(B, T, V, seed=100 + rank, edge_rows=False)
        adv = torch.linspace(-1.0, 1.0, B) * (rank + 1)
        with torch.no_grad():
            lp0 = O.selective_log_softmax(logits.float(), ids)
        gen = torch.Generator().manual_seed(7 + rank)
        old = lp0 + torch.randn(B, T, generator=gen) * 0.3
        ref = lp0 + torch.randn(B, T, generator=gen) * 0.1
        fn = S.GRPOLoss(beta=0.04, loss_type="bnpo", max_completion_length=T)
        o = fn(logits.to(dev), ids.to(dev), mask.to(dev), adv.to(dev), old.to(dev), ref.to(dev))
        g = D.gather_metrics(o.metrics)  # [world, 8] on the host, NCCL
        assert g.shape == (world, 8)
        red = D.reduce_metrics(g)
        # oracle: every rank's local means recomputed on the CPU, reduced as grpo_trainer.py:2150-2172 does
        cfg = O.GRPOConfigLite(beta=0.04, loss_type="bnpo", max_completion_length=T)
        rows = []
        for r in range(world):
            lg, idr, mk = O.synth_batch(B, T, V, seed=100 + r, edge_rows=False)
            ar = torch.linspace(-1.0, 1.0, B) * (r + 1)
            with torch.no_grad():
                l0 = O.selective_log_softmax(lg.float(), idr)
            gr = torch.Generator().manual_seed(7 + r)
            ol = l0 + torch.randn(B, T, generator=gr) * 0.3
            rf = l0 + torch.randn(B, T, generator=gr) * 0.1
            _, met, _, _ = O.grpo_compute_loss(lg.float(), idr, mk, ar, cfg, ol, rf)
            rows.append(met)
        for key, fold in (("kl", "mean"), ("entropy", "mean"), ("clip_ratio/low_mean", "mean"),
                          ("clip_ratio/low_min", "min"), ("clip_ratio/high_mean", "mean"),
                          ("clip_ratio/high_max", "max"), ("clip_ratio/region_mean", "mean")):
            src = {"clip_ratio/low_mean": "clip_ratio/low", "clip_ratio/low_min": "clip_ratio/low",
                   "clip_ratio/high_mean": "clip_ratio/high", "clip_ratio/high_max": "clip_ratio/high",
                   "clip_ratio/region_mean": "clip_ratio/region"}.get(key, key)
            vals = torch.tensor([float(m[src]) for m in rows])
            ref_v = {"mean": vals.nanmean(), "min": vals.min(), "max": vals.max()}[fold].item()
            assert red[key] == pytest.approx(ref_v, rel=1e-4, abs=1e-6), key
        # the deferred ring gives the same rows with ONE exchange for several steps
        ring = torch.stack([o.metrics, o.metrics * 2])
        gr2 = D.gather_metric_rows(ring)
        assert gr2.shape == (world, 2, 8) and torch.equal(gr2[:, 0], g) and torch.equal(gr2[:, 1], g * 2)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, repr(e) + "\n" + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _run(world, case_idx):
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case_idx, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


@pytest.mark.parametrize("case_idx", [1, 2, 3, 5])  # aligned groups, two reward functions, straddling groups, NaN reward
def test_nccl_group_advantages_world2(case_idx):
    _run(2, case_idx)


def test_nccl_group_advantages_world8():
    _run(8, 6)  # config-5 geometry: 256 sequences, 8 completions per prompt, 8 ranks
