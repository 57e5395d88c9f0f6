"""GraphedStep (swh-trl_b200/graphs.py): a forward + backward of the hot path captured in a CUDA graph must give, on
NEW inputs copied into its static buffers, exactly what the eager call gives on them -- same kernels, same launch
geometry, so bit for bit for the log-probs / dlogits, and to the last bits of the atomically summed loss."""
import pytest
import torch

from oracle import trl_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def S():
    import swh_trl_b200 as s
    return s


def _grpo_inputs(B, T, V, seed):
    logits, ids, mask = O.synth_batch(B, T, V, seed=seed, edge_rows=False)
    g = torch.Generator().manual_seed(seed)
    adv = torch.randn(B, generator=g)
    old = -torch.rand(B, T, generator=g) * 3.0
    return {"logits": logits.to(DEV), "ids": ids.to(DEV), "mask": mask.to(DEV), "adv": adv.to(DEV), "old": old.to(DEV)}


@pytest.mark.parametrize("level,V", [("token", 32000), ("sequence", 32000), ("token", 50257)])
def test_graphed_grpo_step_equals_eager(S, level, V):
    """configs[0] geometry (B=4, T=256); 'sequence' takes the two-phase schedule (three launches in the graph),
    V = 50257 the skewed-row consumer."""
    B, T = 4, 256
    fn = S.GRPOLoss(beta=0.0, importance_sampling_level=level, max_completion_length=T)
    a, b = _grpo_inputs(B, T, V, 3), _grpo_inputs(B, T, V, 4)

    def eager(inp):
        x = inp["logits"].clone().requires_grad_(True)
        out = fn(x, inp["ids"], inp["mask"], inp["adv"], inp["old"])
        out.loss.backward()
        return out.loss.detach().clone(), out.per_token_logps.detach().clone(), x.grad.clone(), out.metrics.clone()

    static = {k: v.clone() for k, v in a.items()}
    static["logits"].requires_grad_(True)

    def body(s):
        s["logits"].grad = None
        out = fn(s["logits"], s["ids"], s["mask"], s["adv"], s["old"])
        out.loss.backward()
        return {"loss": out.loss, "logp": out.per_token_logps, "dlogits": s["logits"].grad, "metrics": out.metrics}

    step = S.GraphedStep(body, static)
    for inp in (b, a, b):  # the buffers are overwritten by every replay: go back and forth
        want_loss, want_lp, want_grad, want_metrics = eager(inp)
        res = step.replay(**inp)
        torch.cuda.synchronize()
        assert torch.equal(res["logp"].detach(), want_lp)
        assert torch.equal(res["dlogits"], want_grad)
        torch.testing.assert_close(res["loss"].detach(), want_loss, rtol=1e-6, atol=1e-9)
        torch.testing.assert_close(res["metrics"].detach(), want_metrics, rtol=1e-6, atol=1e-9)
    assert step.replays == 3


def test_graphed_ppo_microbatch_step_equals_eager(S):
    """ppo_trainer.py:557-605 at a cut-down config 3 (mb=4, T=64, V=50304)."""
    mb, T, V = 4, 64, 50304

    def make(seed):
        g = torch.Generator().manual_seed(seed)
        return {"logits": torch.randn(mb, T, V, generator=g).to(torch.bfloat16).to(DEV),
                "vpred": torch.randn(mb, T, generator=g).to(DEV),
                "resp": torch.randint(0, V, (mb, T), generator=g).to(DEV),
                "lp": (-torch.rand(mb, T, generator=g) * 4).to(DEV),
                "adv": torch.randn(mb, T, generator=g).to(DEV),
                "ret": torch.randn(mb, T, generator=g).to(DEV),
                "val": torch.randn(mb, T, generator=g).to(DEV),
                "len": torch.randint(T // 2, T, (mb,), generator=g).to(DEV)}

    def run(s, x, v):
        out = S.ppo_loss(x, s["resp"], s["lp"], s["adv"], s["ret"], s["val"], v, s["len"])
        out.loss.backward()
        return out

    a, b = make(5), make(6)
    static = {k: t.clone() for k, t in a.items()}
    static["logits"].requires_grad_(True)
    static["vpred"].requires_grad_(True)

    def body(s):
        s["logits"].grad = None
        s["vpred"].grad = None
        out = run(s, s["logits"], s["vpred"])
        return {"loss": out.loss, "stats": out.stats, "dlogits": s["logits"].grad, "dvpred": s["vpred"].grad}

    step = S.GraphedStep(body, static)
    x = b["logits"].clone().requires_grad_(True)
    v = b["vpred"].clone().requires_grad_(True)
    want = run(b, x, v)
    res = step(**b)
    torch.cuda.synchronize()
    assert torch.equal(res["dlogits"], x.grad)
    torch.testing.assert_close(res["dvpred"], v.grad, rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(res["loss"].detach(), want.loss.detach(), rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(res["stats"].detach(), want.stats.detach(), rtol=1e-6, atol=1e-9)


def test_graphed_step_rejects_other_shapes(S):
    x = torch.zeros(2, 8, 32000, dtype=torch.bfloat16, device=DEV)
    ids = torch.zeros(2, 8, dtype=torch.long, device=DEV)
    step = S.GraphedStep(lambda s: {"logp": S.logprobs_and_entropy(s["x"], s["ids"], 1.0)[0]}, {"x": x, "ids": ids})
    with pytest.raises(ValueError):
        step.replay(x=torch.zeros(2, 9, 32000, dtype=torch.bfloat16, device=DEV))
    with pytest.raises(ValueError):
        step.replay(x=torch.zeros(2, 8, 32000, dtype=torch.float16, device=DEV))
    with pytest.raises(KeyError):
        step.replay(nope=x)


def test_graphed_step_owns_its_workspaces(S):
    """The accumulators a captured step uses are its own: an eager call on the same stream that needs (and
    re-allocates) a larger shared workspace afterwards must not disturb the graph."""
    B, T, V = 4, 64, 32000
    fn = S.GRPOLoss(beta=0.0, max_completion_length=T)
    a = _grpo_inputs(B, T, V, 8)
    static = {k: v.clone() for k, v in a.items()}
    static["logits"].requires_grad_(True)

    def body(s):
        s["logits"].grad = None
        out = fn(s["logits"], s["ids"], s["mask"], s["adv"], s["old"])
        out.loss.backward()
        return {"loss": out.loss, "dlogits": s["logits"].grad}

    step = S.GraphedStep(body, static)
    assert step._workspaces, "the fused step's accumulators should have been taken from the step's own store"
    first = {k: v.detach().clone() for k, v in step.replay().items()}
    big = _grpo_inputs(96, 16, V, 9)  # more sequences -> a larger shared workspace is allocated by this eager call
    x = big["logits"].requires_grad_(True)
    S.GRPOLoss(beta=0.0, max_completion_length=16)(x, big["ids"], big["mask"], big["adv"], big["old"]).loss.backward()
    again = step.replay()
    torch.cuda.synchronize()
    assert torch.equal(again["dlogits"], first["dlogits"])
    torch.testing.assert_close(again["loss"].detach(), first["loss"], rtol=1e-6, atol=1e-9)
