"""world_size-2 gloo tests of the N>1 host path (no GPU): rank-major reward gather + process_slice, and the
packed metric exchange, against the reference's per-rank outputs (tests/golden/advantages.pt).  The K3 kernel
itself is GPU-only; here the checker (oracle) stands in for it so the *exchange and slicing logic* is what is
under test."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import trl_oracle as O
from tests.conftest import load_golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case_idx, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from swh_trl_b200 import distributed as D
        c = load_golden("advantages.pt")[case_idx]
        n_local = c["B_global"] // world
        local = c["rewards_per_func"][rank * n_local:(rank + 1) * n_local].clone()
        full = D.gather_rewards(local)
        assert torch.equal(full.nan_to_num(7.0), c["rewards_per_func"].nan_to_num(7.0))  # rank-major, bit-exact
        off, cnt = D.process_slice(n_local)
        assert (off, cnt) == (rank * n_local, n_local)
        adv_all = O.group_advantages(full, c["weights"], c["G"], c["scale_rewards"])[1]
        mine = adv_all[off:off + cnt]
        assert torch.equal(mine, c["per_rank"][rank]["advantages"])
        # packed metrics: reference logs gather(x).nanmean() etc. of per-rank local means (grpo_trainer.py:2150-2172)
        m = torch.tensor([0.0, 0.1 * (rank + 1), 2.0 + rank, 0.25 * rank, float("nan") if rank else 0.5, 0.3, 10, 0])
        g = D.gather_metrics(m)
        assert g.shape == (world, 8)
        red = D.reduce_metrics(g)
        want_kl = torch.tensor([0.1 * (r + 1) for r in range(world)]).mean().item()
        assert red["kl"] == pytest.approx(want_kl)
        assert red["clip_ratio/low_min"] == 0.0
        assert red["clip_ratio/high_max"] == 0.5 and red["clip_ratio/high_mean"] == 0.5  # nan-aware
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case_idx", [1, 2, 5])  # world==2 cases: aligned groups, NaN reward, straddling group
def test_gather_and_slice_world2(case_idx):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case_idx, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
