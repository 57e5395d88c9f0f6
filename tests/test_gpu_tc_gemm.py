"""GPU parity of the CTA-pair tcgen05 GEMM family (K7) and of the paths built on it.

The oracle of a floating-point contraction is the plain fp32 / fp64 product of the same bf16 operands on the CPU;
bf16 outputs must equal that product rounded to bf16 up to one ulp (the accumulation order inside the tensor core
differs), fp32 accumulations must agree to a few fp32 ulps of the largest partial sum: the tensor core aligns the
products of one MMA to their largest exponent before adding, so a result that cancels to ~0 carries an absolute error
of about 2^-23 x |largest partial sum| per MMA step, i.e. ~8e-6 sqrt(K) for unit-scale operands (written as atol).
"""
import pytest
import torch

from oracle import trl_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

DEV = "cuda:0"
BF16_ULP = 2.0 ** -7


@pytest.fixture(scope="module")
def S():
    import swh_trl_b200 as s
    return s


def _operands(M, N, K, seed):
    g = torch.Generator().manual_seed(seed)
    a = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16)
    b = (torch.randn(N, K, generator=g) * 0.5).to(torch.bfloat16)
    return a, b


SHAPES = [(256, 256, 64), (128, 256, 128), (1, 8, 8), (300, 520, 200), (512, 768, 1024), (700, 1000, 4096),
          (2048, 3584, 512)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_gemm_kmajor_store(S, M, N, K):
    """logits_c = hidden_c W^T: both operands K-major, bf16 output, optional bias; ragged tiles in M, N and K."""
    from swh_trl_b200 import ops
    a, b = _operands(M, N, K, M + N + K)
    want = a.double() @ b.double().t()
    got = ops.tc_gemm(a.to(DEV), b.to(DEV))
    torch.testing.assert_close(got.float().cpu(), want.to(torch.bfloat16).float(), rtol=BF16_ULP, atol=8e-6 * K ** 0.5)
    bias = torch.linspace(-1, 1, N).to(torch.bfloat16)
    got_b = ops.tc_gemm(a.to(DEV), b.to(DEV), bias=bias.to(DEV))
    want_b = (want + bias.double()).to(torch.bfloat16).float()
    torch.testing.assert_close(got_b.float().cpu(), want_b, rtol=BF16_ULP, atol=8e-6 * K ** 0.5)
    # strided operands and output (row pitch > row length)
    if K % 8 == 0 and N % 8 == 0:
        ap = torch.zeros(M, K + 16, dtype=torch.bfloat16, device=DEV)
        ap[:, :K] = a.to(DEV)
        out = torch.full((M, N + 8), 7.0, dtype=torch.bfloat16, device=DEV)
        ops.tc_gemm(ap[:, :K], b.to(DEV), out=out[:, :N])
        assert torch.equal(out[:, :N], got) and bool((out[:, N:] == 7).all())


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_gemm_b_mnmajor_store(S, M, N, K):
    """dH_c = dlogits_c W: A K-major, B stored [K, N] (MN-major), bf16 output."""
    from swh_trl_b200 import ops
    if N % 8:
        pytest.skip("an MN-major operand needs 16-byte rows")
    a, b = _operands(M, N, K, 3 * M + N + K)
    want = (a.double() @ b.double().t()).to(torch.bfloat16).float()
    got = ops.tc_gemm(a.to(DEV), b.t().contiguous().to(DEV), b_layout=1)
    torch.testing.assert_close(got.float().cpu(), want, rtol=BF16_ULP, atol=8e-6 * K ** 0.5)


@pytest.mark.parametrize("b_layout", [0, 1])
def test_tc_gemm_split_k(S, b_layout):
    """Few output tiles and a long contraction: K is split into slices whose fp32 partial products are added in slice
    order.  Same result as the unsplit kernel up to the bf16 rounding of a differently ordered fp32 sum, and
    bit-identical from run to run."""
    from swh_trl_b200 import ops
    M, N, K = 512, 512, 16384
    a, b = _operands(M, N, K, 99)
    bd = (b.t().contiguous() if b_layout else b).to(DEV)
    want = (a.double() @ b.double().t()).to(torch.bfloat16).float()
    one = ops.tc_gemm(a.to(DEV), bd, b_layout=b_layout, split_k=False)
    two = ops.tc_gemm(a.to(DEV), bd, b_layout=b_layout, split_k=True)
    again = ops.tc_gemm(a.to(DEV), bd, b_layout=b_layout, split_k=True)
    assert torch.equal(two, again)
    for got in (one, two):
        torch.testing.assert_close(got.float().cpu(), want, rtol=BF16_ULP, atol=8e-6 * K ** 0.5)


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_gemm_mnmajor_accumulate(S, M, N, K):
    """dW += dlogits_c^T hidden_c: both operands stored [K, rows], fp32 accumulation into the output, twice."""
    from swh_trl_b200 import ops
    if M % 8 or N % 8:
        pytest.skip("an MN-major operand needs 16-byte rows")
    a, b = _operands(M, N, K, 5 * M + N + K)
    want = a.double() @ b.double().t()
    out = torch.full((M, N), 0.25, dtype=torch.float32, device=DEV)
    ops.tc_gemm(a.t().contiguous().to(DEV), b.t().contiguous().to(DEV), a_layout=1, b_layout=1, out=out, accumulate=True,
                m_fastest=False)
    tol = 8e-6 * K ** 0.5
    torch.testing.assert_close(out.cpu().double(), want + 0.25, rtol=1e-5, atol=tol)
    ops.tc_gemm(a.t().contiguous().to(DEV), b.t().contiguous().to(DEV), a_layout=1, b_layout=1, out=out, accumulate=True)
    torch.testing.assert_close(out.cpu().double(), 2 * want + 0.25, rtol=1e-5, atol=2 * tol)
    # first chunk of dW: plain fp32 store; last chunk: bf16(accumulator + D) rounded in the epilogue
    at, bt = a.t().contiguous().to(DEV), b.t().contiguous().to(DEV)
    first = ops.tc_gemm(at, bt, a_layout=1, b_layout=1, out_fp32=True, m_fastest=False)
    torch.testing.assert_close(first.cpu().double(), want, rtol=1e-5, atol=tol)
    final = ops.tc_gemm(at, bt, a_layout=1, b_layout=1, addend=first, m_fastest=False)
    assert final.dtype == torch.bfloat16
    torch.testing.assert_close(final.float().cpu(), (2 * want).to(torch.bfloat16).float(), rtol=BF16_ULP, atol=2 * tol)


def test_fused_forward_pair_vs_single_cta(S):
    """The CTA-pair statistics kernel and the single-CTA kernel agree (same tile order, fp32 in TMEM both)."""
    import subprocess
    import sys
    code = (
        "import sys, torch, swh_trl_b200 as S\n"
        "g = torch.Generator().manual_seed(3)\n"
        "h = torch.randn(1000, 512, generator=g).to(torch.bfloat16).cuda()\n"
        "W = (torch.randn(30000, 512, generator=g) * 0.08).to(torch.bfloat16).cuda()\n"
        "ids = torch.randint(0, 30000, (1000,), generator=g).cuda()\n"
        "lp, ent = S.fused_linear_logprobs(h, W, ids, temperature=0.9)\n"
        "torch.save((lp.cpu(), ent.cpu()), sys.argv[1])\n"
    )
    import os
    import tempfile
    outs = []
    for impl in ("1", "2"):
        with tempfile.NamedTemporaryFile(suffix=".pt") as f:
            env = dict(os.environ, B200TRL_K5_IMPL=impl)
            subprocess.run([sys.executable, "-c", code, f.name], check=True, env=env, timeout=300,
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            outs.append(torch.load(f.name))
    torch.testing.assert_close(outs[0][0], outs[1][0], rtol=0, atol=2e-6)
    torch.testing.assert_close(outs[0][1], outs[1][1], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("mask", [1, 2, 4, 7])
def test_seam_on_tcgen05_gemms(S, mask):
    """The Liger-shaped operator with its GEMMs on the K7 kernel (bit 0 logits, 1 dH, 2 dW) against the same operator
    on the library GEMMs: the bf16 logits may differ by one ulp where the accumulation order flips a rounding, so loss
    and gradients agree to the bf16 level, and against the fp32 oracle of the reference's non-Liger path."""
    from swh_trl_b200 import ops
    B, T, H, V = 4, 128, 256, 4096
    g = torch.Generator().manual_seed(11)
    hidden = torch.randn(B, T, H, generator=g).to(torch.bfloat16).to(DEV)
    W = (torch.randn(V, H, generator=g) * 0.06).to(torch.bfloat16).to(DEV)
    ids = torch.randint(0, V, (B, T), generator=g).to(DEV)
    cmask = torch.ones(B, T, dtype=torch.int32, device=DEV)
    cmask[1, 90:] = 0
    adv = torch.tensor([0.5, -1.5, 1.0, -0.25], device=DEV)
    cfg = ops.make_cfg(0.0, 0.2, 0.2, None, "bnpo", "token", T)
    prev = ops.set_seam_gemm_mask(0)
    try:
        base = ops.fused_linear_grpo(hidden, W, None, ids, cmask, adv, None, None, cfg, 1.0, 2, True, True, False)
        ops.set_seam_gemm_mask(mask)
        ours = ops.fused_linear_grpo(hidden, W, None, ids, cmask, adv, None, None, cfg, 1.0, 2, True, True, False)
    finally:
        ops.set_seam_gemm_mask(prev)
    torch.cuda.synchronize()
    loss0, _, lp0, _, dh0, dw0, _ = base
    loss1, _, lp1, _, dh1, dw1, _ = ours
    assert loss1.item() == pytest.approx(loss0.item(), rel=1e-3, abs=1e-6)
    torch.testing.assert_close(lp1, lp0, rtol=0, atol=2e-2)  # one bf16 ulp of a logit of magnitude ~2
    assert (dh1.float() - dh0.float()).norm() <= 1e-2 * dh0.float().norm()
    assert (dw1.float() - dw0.float()).norm() <= 1e-2 * dw0.float().norm()
    # the oracle: reference loss on bf16-rounded logits of the same operands (what a bf16 model hands the loss)
    logits = (hidden.float().cpu() @ W.float().cpu().t()).to(torch.bfloat16).float()
    ocfg = O.GRPOConfigLite(beta=0.0, loss_type="bnpo", max_completion_length=T)
    want, _, _, _ = O.grpo_compute_loss(logits, ids.cpu(), cmask.cpu(), adv.cpu(), ocfg, None, None)
    assert loss1.item() == pytest.approx(want.item(), rel=1e-3, abs=1e-6)
