#!/usr/bin/env python
"""Randomised check of the CTA-pair tcgen05 GEMM family (K7): random M / N / K (ragged in every dimension), operand
layouts, epilogues (bf16 store +/- bias, fp32 store, fp32 accumulate, bf16 store with fp32 addend, split-K) against the
fp64 product of the same bf16 operands.  Bars of tests/test_gpu_tc_gemm.py: bf16 outputs one ulp, fp32 outputs
within the accumulator's truncation bound (see `atol`).  One JSON line; exit code 1 on the first disagreement.  TF_CASES (default 80), TF_SEED."""
import json, os, random, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swh_trl_b200 as S  # noqa: E402,F401
from swh_trl_b200 import ops  # noqa: E402

DEV = torch.device("cuda", 0)
N_CASES, SEED = int(os.environ.get("TF_CASES", 80)), int(os.environ.get("TF_SEED", 0))
rnd = random.Random(SEED)
BF16_ULP = 2.0 ** -7
worst = {"bf16_ulp": 0.0, "f32_abs_over_bar": 0.0}
for case in range(N_CASES):
    M = rnd.choice([rnd.randint(1, 700), rnd.randint(700, 5000), rnd.choice([128, 256, 512, 4096])])
    N = rnd.choice([rnd.randint(1, 900), rnd.randint(900, 6000), rnd.choice([256, 3584, 1024])])
    K = rnd.choice([rnd.randint(8, 600), rnd.randint(600, 9000), rnd.choice([64, 3584, 4096])])
    K -= K % 8  # 16-byte rows for the K-major tensor maps
    K = max(K, 8)
    mode = rnd.choice(["bf16", "bf16_bias", "f32", "acc", "addend", "bf16_nosplit"])
    # the instantiated (A layout, B layout) x epilogue combinations are the ones the seam uses: bf16 stores for
    # (K, K) logits, (K, MN) dH, (MN, MN) dW; fp32 store / accumulate for (MN, MN) dW
    a_l, b_l = rnd.choice([(0, 0), (0, 1), (1, 1)]) if mode.startswith("bf16") or mode == "addend" else (1, 1)
    if a_l:
        M = max(8, M - M % 8)
    if b_l:
        N = max(8, N - N % 8)
    if mode in ("bf16", "bf16_bias", "bf16_nosplit", "addend"):
        N = max(8, N - N % 8)  # bf16 output rows are 16-byte multiples for the TMA store
    if mode in ("f32", "acc"):
        N = max(4, N - N % 4)
    g = torch.Generator(device=DEV).manual_seed(SEED * 7919 + case)
    a = (torch.randn(M, K, generator=g, device=DEV) * 0.5).to(torch.bfloat16)
    b = (torch.randn(N, K, generator=g, device=DEV) * 0.5).to(torch.bfloat16)
    want = a.double() @ b.double().t()
    A = a.t().contiguous() if a_l else a
    Bm = b.t().contiguous() if b_l else b
    desc = dict(case=case, M=M, N=N, K=K, a_layout=a_l, b_layout=b_l, mode=mode)
    try:
        # the tensor core aligns the products of an MMA step to the running sum's exponent and truncates: up to half an
        # fp32 ulp of the accumulator per 16-deep step, same sign -> (K / 16) * 2^-24 * |largest sum| at worst
        atol = max(8e-6 * K ** 0.5, (K / 16) * 2.0 ** -24 * float(want.abs().max()))
        if mode in ("bf16", "bf16_nosplit"):
            got = ops.tc_gemm(A, Bm, a_l, b_l, split_k=(mode == "bf16"), m_fastest=bool(rnd.randint(0, 1)))
            ref = want.to(torch.bfloat16).float()
            ok = torch.isclose(got.float(), ref, rtol=BF16_ULP, atol=atol).all()
        elif mode == "bf16_bias":
            bias = torch.linspace(-1, 1, N, device=DEV).to(torch.bfloat16)
            got = ops.tc_gemm(A, Bm, a_l, b_l, bias=bias)
            ref = (want + bias.double()).to(torch.bfloat16).float()
            ok = torch.isclose(got.float(), ref, rtol=BF16_ULP, atol=atol).all()
        elif mode == "f32":
            got = ops.tc_gemm(A, Bm, a_l, b_l, out_fp32=True)
            ok = torch.isclose(got.double(), want, rtol=2e-6, atol=atol).all()
            if not bool(ok):  # how far is the library's tensor-core GEMM with the same operands?
                lib_err = float((torch.matmul(a, b.t()).double() - want).abs().max())  # bf16 out: rounding dominates
                lib32 = float((torch.matmul(a.float(), b.float().t()).double() - want).abs().max())
                desc.update(max_abs_err=float((got.double() - want).abs().max()), bar=atol, torch_bf16_out_err=lib_err,
                            torch_fp32_gemm_err=lib32, want_absmax=float(want.abs().max()))
        elif mode == "acc":
            base = torch.randn(M, N, generator=g, device=DEV)
            out = base.clone()
            ops.tc_gemm(A, Bm, a_l, b_l, out=out, accumulate=True)
            ops.tc_gemm(A, Bm, a_l, b_l, out=out, accumulate=True)
            ok = torch.isclose(out.double(), base.double() + 2 * want, rtol=4e-6, atol=2 * atol).all()
        else:
            add = torch.randn(M, N, generator=g, device=DEV)
            got = ops.tc_gemm(A, Bm, a_l, b_l, addend=add)
            ref = (want + add.double()).to(torch.bfloat16).float()
            ok = torch.isclose(got.float(), ref, rtol=BF16_ULP, atol=atol).all()
        torch.cuda.synchronize()
        assert bool(ok), mode
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"failed": repr(e)[:300], **desc}))
        sys.exit(1)
print(json.dumps({"cases": N_CASES, "seed": SEED, "agree": True}))
