"""Kernel-level numerics on the B200: every C-ABI kernel family against a plain PyTorch
fp32 evaluation of the same op on the same (bf16-rounded) inputs."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from eilev_b200 import ops
    return ops


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


def _close(got, ref, atol, rtol, what=""):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = (err > tol).sum().item()
    assert bad == 0, f"{what}: {bad}/{err.numel()} off, max err {err.max().item():.4g} (ref max {ref.abs().max().item():.4g})"


GEMM_SHAPES = [
    # (M, N, K, block_n)
    (128, 256, 64, 256), (128, 256, 128, 256), (256, 512, 256, 256), (128, 128, 64, 128),
    (128, 176, 64, 176), (128, 64, 64, 64), (300, 1408, 1408, 0), (257, 4224, 1408, 0),
    (1000, 6144, 1408, 0), (520, 1408, 6144, 0), (544, 768, 768, 0), (976, 2560, 2560, 0),
    (200, 264, 72, 0), (64, 50272, 256, 0), (4096, 1408, 592, 0),
    # CTA-pair kernel (block_n = 1000 + BN): M tails, a fully out-of-range peer CTA, K tails
    (512, 512, 256, 1256), (300, 1408, 1408, 1176), (257, 4224, 1408, 1256), (1000, 6144, 1408, 1176),
    (520, 1408, 6144, 1256), (200, 264, 72, 1128), (100, 256, 64, 1256), (8200, 1408, 1408, 0),
    (5000, 6144, 1408, 0), (4100, 768, 1408, 1128),
    # run-time tile widths of the CTA-pair kernel (wave-filling widths of the OPT shapes at M = 976): widths that
    # are not a multiple of 64 end in a partial slab written with direct stores
    (976, 7680, 2560, 0), (976, 10240, 512, 0), (976, 2560, 512, 1144), (976, 1000, 256, 1208),
    (600, 2048, 128, 1192), (976, 2560, 320, 1096), (1500, 1408, 192, 1240), (976, 2560, 192, 1032),
]


@pytest.mark.parametrize("m,n,k,bn", GEMM_SHAPES)
def test_gemm_tcgen05_plain(m, n, k, bn):
    ops = _ops()
    a, w = _rand(m, k, seed=1), _rand(n, k, seed=2)
    out = ops.gemm(a, w, backend=ops.GEMM_TCGEN05, block_n=bn)
    ref = a.float() @ w.float().t()
    _close(out, ref, atol=0.02 * math.sqrt(k), rtol=0.01, what=f"gemm {m}x{n}x{k} bn={bn}")


@pytest.mark.parametrize("tokens,n_out,n_in", [(544, 768, 768), (544, 3072, 768), (544, 768, 3072), (100, 264, 72),
                                               (1000, 128, 64), (37, 768, 1408), (544, 2560, 768)])
def test_gemm_weight_gradient_from_untransposed_activations(tokens, n_out, n_in):
    """gemm_tn: dW = dY^T X with both operands entering the instruction MN-major (operand_layout = 1), against
    fp32 torch and against the transpose-kernel path; accumulation into an f32 buffer (beta = 1)."""
    ops = _ops()
    dy, x = _rand(tokens, n_out, scale=0.3, seed=71), _rand(tokens, n_in, scale=0.5, seed=72)
    ref = dy.float().t() @ x.float()
    got = ops.gemm_tn(dy, x)
    assert got.dtype == torch.float32
    _close(got, ref, atol=0.02 * math.sqrt(tokens), rtol=0.01, what=f"gemm_tn {tokens}x{n_out}x{n_in}")
    two = ops.gemm(ops.transpose(dy), ops.transpose(x), out_dtype=torch.float32)
    _close(got, two, atol=1e-3 * math.sqrt(tokens), rtol=1e-3, what="gemm_tn vs transposes")
    acc = torch.ones(n_out, n_in, device="cuda")
    ops.gemm_tn(dy, x, out=acc, beta=1.0)
    _close(acc, ref + 1.0, atol=0.02 * math.sqrt(tokens), rtol=0.01, what="gemm_tn accumulate")
    # column slices of a wider buffer (row stride != width)
    wide = _rand(tokens, n_out + 64, scale=0.3, seed=73)
    got2 = ops.gemm_tn(wide[:, 64:], x)
    _close(got2, wide[:, 64:].float().t() @ x.float(), atol=0.02 * math.sqrt(tokens), rtol=0.01, what="gemm_tn strided")


@pytest.mark.parametrize("bn", [1144, 1208, 1192, 0])
def test_gemm_cta_pair_widths_with_full_epilogue(bn):
    """bias + ReLU + dropout + residual through a tile width with a partial last slab (staged slabs and direct
    chunks in one tile) at the OPT row count."""
    ops = _ops()
    m, n, k = 976, 2560, 384
    a, w = _rand(m, k, scale=0.5, seed=13), _rand(n, k, scale=0.1, seed=14)
    bias = torch.randn(n, device="cuda")
    res = _rand(m, n, seed=15)
    seed = torch.tensor([77], dtype=torch.int64, device="cuda")
    out = ops.gemm(a, w, bias, residual=res, epilogue=ops.EPI_RELU, dropout=(0.25, seed, 5),
                   backend=ops.GEMM_TCGEN05, block_n=bn)
    keep = ops.dropout(torch.ones(m, n, dtype=torch.bfloat16, device="cuda"), 0.25, seed, 5).float()
    ref = torch.relu(a.float() @ w.float().t() + bias) * keep + res.float()
    _close(out, ref, atol=0.04, rtol=0.01, what=f"cta-pair width {bn}")


@pytest.mark.parametrize("backend", ["tcgen05", "tcgen05_2cta", "generic"])
@pytest.mark.parametrize("epi", ["none", "gelu", "relu"])
def test_gemm_epilogues(backend, epi):
    ops = _ops()
    m, n, k = 384, 704, 320
    a, w = _rand(m, k, scale=0.5, seed=3), _rand(n, k, scale=0.1, seed=4)
    bias = torch.randn(n, device="cuda")
    res = _rand(m, n, seed=5)
    e = {"none": ops.EPI_NONE, "gelu": ops.EPI_GELU, "relu": ops.EPI_RELU}[epi]
    be = ops.GEMM_GENERIC if backend == "generic" else ops.GEMM_TCGEN05
    bn = 1176 if backend == "tcgen05_2cta" else 0
    out = ops.gemm(a, w, bias, residual=res, epilogue=e, alpha=0.5, alpha_cols=352, backend=be, block_n=bn)
    pre = a.float() @ w.float().t() + bias
    pre[:, :352] *= 0.5
    act = {"none": lambda x: x, "gelu": torch.nn.functional.gelu, "relu": torch.relu}[epi](pre)
    _close(out, act + res.float(), atol=0.03, rtol=0.01, what=f"{backend}/{epi}")
    # f32 output with accumulation
    c = torch.ones(m, n, device="cuda")
    ops.gemm(a, w, bias, out=c, beta=1.0, backend=be, block_n=bn)
    _close(c, a.float() @ w.float().t() + bias + 1.0, atol=0.03, rtol=0.01, what="f32 beta")


@pytest.mark.parametrize("m,n,k,bn,epi", [(1000, 4224, 1408, 0, "none"), (5000, 6144, 1408, 0, "gelu"),
                                           (300, 704, 320, 1256, "gelu"), (257, 264, 72, 0, "none"),
                                           (640, 1408, 1408, 1176, "relu"), (130, 24, 8, 0, "none")])
def test_gemm_layernorm_fold(m, n, k, bn, epi):
    """LayerNorm folded into the consuming GEMM (vb_gemm_args.ln_stats): act(LN(x) W^T + b) from the
    un-normalised x, W*gamma, b + W beta, the row statistics and colsum(W*gamma).  The rows carry a
    mean of the size of their spread so the -rstd*mean*colsum term matters."""
    ops = _ops()
    from eilev_b200.engine.packing import ln_fold
    g = torch.Generator(device="cuda").manual_seed(21)
    x = (torch.randn(m, k, generator=g, device="cuda") * (0.5 + torch.rand(m, 1, generator=g, device="cuda") * 3)
         + torch.randn(m, 1, generator=g, device="cuda") * 2).to(torch.bfloat16)
    w = _rand(n, k, scale=0.05, seed=22)
    b = torch.randn(n, device="cuda") * 0.1
    gamma = 1.0 + 0.3 * torch.randn(k, device="cuda")
    beta = 0.2 * torch.randn(k, device="cuda")
    wg, bias, cs = ln_fold(w, b, gamma, beta)
    st = ops.row_stats(x)
    assert st.dtype == torch.float64
    assert torch.allclose(st[:, 0], x.double().sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(st[:, 1], x.double().pow(2).sum(1), rtol=1e-4, atol=1e-2)
    e = {"none": ops.EPI_NONE, "gelu": ops.EPI_GELU, "relu": ops.EPI_RELU}[epi]
    out = ops.gemm(x, wg, bias, epilogue=e, ln_fold=(st, cs, 1e-6), backend=ops.GEMM_TCGEN05, block_n=bn)
    pre = torch.nn.functional.layer_norm(x.float(), (k,), gamma, beta, 1e-6) @ w.float().t() + b
    ref = {"none": lambda t: t, "gelu": torch.nn.functional.gelu, "relu": torch.relu}[epi](pre)
    # the stand-alone path rounds LN(x) to bf16 before the GEMM: same error budget as test_gemm_epilogues
    _close(out, ref, atol=0.03 + 0.004 * math.sqrt(k) * 0.05 * 8, rtol=0.01, what=f"ln fold {m}x{n}x{k} {epi}")
    y = ops.layernorm(x, gamma, beta, 1e-6)
    two_step = ops.gemm(y, w, b, epilogue=e, backend=ops.GEMM_TCGEN05, block_n=bn)
    err_fold = (out.float() - ref).pow(2).mean().sqrt().item()
    err_two = (two_step.float() - ref).pow(2).mean().sqrt().item()
    assert err_fold <= 1.5 * err_two + 1e-3, (err_fold, err_two)  # no less accurate than LayerNorm kernel + GEMM


@pytest.mark.parametrize("m,n,k,bn", [(1000, 1408, 1408, 0), (5000, 1408, 6144, 0), (300, 704, 320, 1256),
                                      (300, 704, 320, 1176), (257, 264, 72, 0), (640, 1408, 1408, 256)])
def test_gemm_output_row_statistics(m, n, k, bn):
    """stats_out: the epilogue adds [sum, sum of squares] of the bf16 values it stores (residual GEMM in
    place, as in the ViT layer); stats_zero: the other buffer is cleared by the same launch."""
    ops = _ops()
    a, w = _rand(m, k, scale=0.5, seed=31), _rand(n, k, scale=0.05, seed=32)
    bias = torch.randn(n, device="cuda") * 0.1
    x = _rand(m, n, seed=33)
    ref = (a.float() @ w.float().t() + bias + x.float()).to(torch.bfloat16)
    st = torch.zeros(m, 2, device="cuda", dtype=torch.float64)
    other = torch.full((m, 2), 7.0, device="cuda", dtype=torch.float64)
    ops.gemm(a, w, bias, residual=x, out=x, stats_out=st, stats_zero=other, backend=ops.GEMM_TCGEN05, block_n=bn)
    _close(x, ref, atol=0.03, rtol=0.01, what="residual gemm")
    got = x.double()
    assert torch.allclose(st[:, 0], got.sum(1), rtol=1e-4, atol=2e-2), (st[:, 0] - got.sum(1)).abs().max()
    assert torch.allclose(st[:, 1], got.pow(2).sum(1), rtol=1e-4, atol=2e-2), (st[:, 1] - got.pow(2).sum(1)).abs().max()
    assert float(other.abs().max()) == 0.0
    # a second launch accumulates on top (the caller owns the zeroing)
    y = _rand(m, n, seed=34)
    ops.gemm(a, w, bias, residual=y, out=y, stats_out=st, backend=ops.GEMM_TCGEN05, block_n=bn)
    assert torch.allclose(st[:, 0], got.sum(1) + y.double().sum(1), rtol=1e-4, atol=4e-2)
    # the statistics are order-independent (f64 sums of f32 partials): a repeat gives the same bits
    x2 = _rand(m, n, seed=33)
    st2 = torch.zeros(m, 2, device="cuda", dtype=torch.float64)
    ops.gemm(a, w, bias, residual=x2, out=x2, stats_out=st2, backend=ops.GEMM_TCGEN05, block_n=bn)
    st3 = torch.zeros(m, 2, device="cuda", dtype=torch.float64)
    x3 = _rand(m, n, seed=33)
    ops.gemm(a, w, bias, residual=x3, out=x3, stats_out=st3, backend=ops.GEMM_TCGEN05, block_n=bn)
    assert torch.equal(st2, st3)


def test_gemm_gelu_epilogue_accuracy():
    """The epilogue's GELU (sigmoid of an odd quintic, common.cuh::gelu_fast) against the exact erf GELU
    the reference uses (HF hidden_act="gelu"): f32 output, inputs spread over [-12, 12]."""
    ops = _ops()
    m, n, k = 256, 512, 64
    a = torch.zeros(m, k, dtype=torch.bfloat16, device="cuda")
    a[:, 0] = 1.0
    w = torch.zeros(n, k, dtype=torch.bfloat16, device="cuda")
    bias = torch.linspace(-12.0, 12.0, n, device="cuda")
    out = torch.empty(m, n, device="cuda")
    ops.gemm(a, w, bias, out=out, epilogue=ops.EPI_GELU, backend=ops.GEMM_TCGEN05)
    ref = torch.nn.functional.gelu(bias.double()).float()
    err = (out[0] - ref).abs().max().item()
    assert err < 6e-5, err
    assert (out - out[0:1]).abs().max().item() == 0.0
    big = torch.tensor([-1e4, -50.0, 50.0, 1e4], device="cuda").repeat(n // 4)
    ops.gemm(a, w, big, out=out, epilogue=ops.EPI_GELU, backend=ops.GEMM_TCGEN05)
    assert torch.equal(out[0], torch.nn.functional.gelu(big))


@pytest.mark.parametrize("backend,bn", [("tcgen05", 0), ("tcgen05", 1256), ("tcgen05", 1176), ("generic", 0)])
@pytest.mark.parametrize("act", ["gelu", "relu"])
def test_gemm_activation_backward_epilogue(backend, bn, act):
    """VB_EPI_{GELU,RELU}_BWD: d_pre = (dy W) * act'(saved) in the dgrad GEMM's epilogue == GEMM then vb_act_bwd."""
    ops = _ops()
    m, n, k = 640, 768, 320
    dy, wt = _rand(m, k, scale=0.5, seed=41), _rand(n, k, scale=0.1, seed=42)
    saved = _rand(m, n, seed=43)
    a = {"gelu": ops.EPI_GELU, "relu": ops.EPI_RELU}[act]
    be = ops.GEMM_GENERIC if backend == "generic" else ops.GEMM_TCGEN05
    e = {"gelu": ops.EPI_GELU_BWD, "relu": ops.EPI_RELU_BWD}[act]
    got = ops.gemm(dy, wt, residual=saved, epilogue=e, backend=be, block_n=bn)
    pre = saved.float().requires_grad_(True)
    y = torch.nn.functional.gelu(pre) if act == "gelu" else torch.relu(pre)
    y.backward(dy.float() @ wt.float().t())
    _close(got, pre.grad, atol=0.03, rtol=0.01, what=f"act bwd epilogue {backend}/{bn}/{act}")
    two = ops.act_bwd(ops.gemm(dy, wt, backend=be, block_n=bn), saved, a)
    _close(got, two, atol=0.03, rtol=0.02, what="fused vs gemm + act_bwd")


def test_gemm_generic_odd_shapes():
    ops = _ops()
    for (m, n, k) in [(5, 7, 3), (65, 24, 192), (130, 8, 8), (33, 100, 50)]:
        a, w = _rand(m, k, seed=6), _rand(n, k, seed=7)
        out = ops.gemm(a, w)
        _close(out, a.float() @ w.float().t(), atol=0.05, rtol=0.01, what=f"generic {m}x{n}x{k}")


def test_gemm_patch_rowgroup():
    ops = _ops()
    frames, p, dim, k = 3, 256, 1408, 640
    a, w = _rand(frames * p, k, scale=0.3, seed=8), _rand(dim, k, scale=0.1, seed=9)
    bias = torch.randn(dim, device="cuda")
    pos = _rand(p + 1, dim, seed=10)
    hidden = torch.zeros(frames, p + 1, dim, dtype=torch.bfloat16, device="cuda")
    ops.gemm(a, w, bias, residual=pos, out=hidden.view(-1, dim), row_group=p)
    ref = (a.float() @ w.float().t() + bias).view(frames, p, dim) + pos[1:].float()
    _close(hidden[:, 1:], ref, atol=0.03, rtol=0.01, what="patch rows")
    assert hidden[:, 0].abs().max().item() == 0.0


@pytest.mark.parametrize("rows,cols", [(1000, 1408), (544, 768), (976, 2560), (37, 8), (64, 100),
                                       (301, 2048), (5, 3072), (3, 2056)])  # >= 2048 columns: one block per row
def test_layernorm_fwd_bwd(rows, cols):
    ops = _ops()
    x, r = _rand(rows, cols, seed=11), _rand(rows, cols, seed=12)
    g = torch.randn(cols, device="cuda")
    b = torch.randn(cols, device="cuda")
    y, mean, rstd = ops.layernorm(x, g, b, 1e-5, residual=r, save_stats=True)
    xin = (x.float() + r.float()).requires_grad_(True)
    gp, bp = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xin, (cols,), gp, bp, 1e-5)
    _close(y, ref, atol=0.03, rtol=0.01, what="ln fwd")
    dy = _rand(rows, cols, seed=13)
    ref.backward(dy.float())
    xin_bf = (x.float() + r.float()).to(torch.bfloat16)
    dgamma = torch.zeros(cols, device="cuda")
    dbeta = torch.zeros(cols, device="cuda")
    dx = ops.layernorm_bwd(dy, xin_bf, g, mean, rstd, dgamma=dgamma, dbeta=dbeta)
    _close(dx, xin.grad, atol=0.05, rtol=0.03, what="ln dx")
    _close(dgamma, gp.grad, atol=0.02 * math.sqrt(rows) + 0.05, rtol=0.03, what="ln dgamma")
    _close(dbeta, bp.grad, atol=0.02 * math.sqrt(rows) + 0.05, rtol=0.03, what="ln dbeta")


def _attn_ref(q, k, v, heads, scale, causal, key_mask):
    b, sq, hd = q.shape
    skv = k.shape[1]
    d = hd // heads
    qh = q.float().view(b, sq, heads, d).transpose(1, 2)
    kh = k.float().view(b, skv, heads, d).transpose(1, 2)
    vh = v.float().view(b, skv, heads, d).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * scale
    if causal:
        i = torch.arange(sq, device=q.device)[:, None]
        j = torch.arange(skv, device=q.device)[None, :]
        s = s.masked_fill(j > i + (skv - sq), -1e30)
    if key_mask is not None:
        s = s.masked_fill(key_mask[:, None, None, :] == 0, -1e30)
    p = torch.softmax(s, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(b, sq, hd)


def _attn_lse_ref(q, k, heads, scale, causal, key_mask):
    """log-sum-exp of the masked, scaled scores, (b, heads, sq); -inf for a row that sees no key."""
    b, sq, hd = q.shape
    skv = k.shape[1]
    d = hd // heads
    qh = q.float().view(b, sq, heads, d).transpose(1, 2)
    kh = k.float().view(b, skv, heads, d).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * scale
    if causal:
        i = torch.arange(sq, device=q.device)[:, None]
        j = torch.arange(skv, device=q.device)[None, :]
        s = s.masked_fill(j > i + (skv - sq), float("-inf"))
    if key_mask is not None:
        s = s.masked_fill(key_mask[:, None, None, :] == 0, float("-inf"))
    return torch.logsumexp(s, dim=-1)


ATTN_CASES = [
    # b, heads, d, sq, skv, causal, masked
    (3, 16, 88, 257, 257, False, False),
    (2, 12, 64, 32, 32, False, False),
    (2, 12, 64, 32, 2056, False, False),
    (1, 3, 64, 40, 700, False, True),      # few queries over a long memory, masked: several key tiles per CTA in bwd
    (2, 2, 88, 64, 1030, False, False),
    (1, 32, 80, 976, 976, True, False),
    (2, 32, 80, 200, 200, True, True),
    (2, 4, 2, 11, 11, True, False),
    (2, 4, 2, 5, 37, False, False),
    # tcgen05 backward: ragged tiles with more than one batch, causal with more keys than queries, d = 128 / 48
    (2, 3, 80, 300, 300, True, False),
    (2, 4, 64, 130, 515, False, True),
    (1, 2, 128, 200, 200, True, False),
    (2, 2, 80, 150, 260, True, False),
    (1, 2, 48, 100, 100, True, True),
    (1, 2, 80, 1500, 1500, True, False),
]


@pytest.mark.parametrize("b,heads,d,sq,skv,causal,masked", ATTN_CASES)
def test_attention_fwd_bwd(b, heads, d, sq, skv, causal, masked):
    ops = _ops()
    hd = heads * d
    scale = d ** -0.5
    if sq == skv:  # fused QKV buffer, as in the ViT / OPT
        qkv = _rand(b, sq, 3 * hd, seed=20)
        q, k, v = qkv[:, :, :hd], qkv[:, :, hd:2 * hd], qkv[:, :, 2 * hd:]
    else:
        q = _rand(b, sq, hd, seed=21)
        kv = _rand(b, skv, 2 * hd, seed=22)
        k, v = kv[:, :, :hd], kv[:, :, hd:]
    key_mask = None
    if masked:
        key_mask = torch.ones(b, skv, dtype=torch.uint8, device="cuda")
        key_mask[0, :17] = 0  # left padding on the first sequence
    o, lse = ops.attention(q, k, v, heads, scale, causal=causal, key_mask=key_mask, need_lse=True)
    # head dims that are a multiple of 16 run the tcgen05 flash kernel (attention_flash_tcgen05.cu)
    want_kernel = "tcgen05_flash" if d % 16 == 0 and sq >= 128 and os.environ.get("VB_ATTN_FWD_TC") != "0" else "mma_sync"
    assert ops.attention_kernel(q, k, v, heads, causal=causal, key_mask=key_mask, need_lse=True) == want_kernel
    lse_ref = _attn_lse_ref(q, k, heads, scale, causal, key_mask)
    fin = torch.isfinite(lse_ref)
    _close(lse[fin], lse_ref[fin], atol=0.02, rtol=0.01, what="attn lse")
    qf = q.float().detach().clone().requires_grad_(True)
    kf = k.float().detach().clone().requires_grad_(True)
    vf = v.float().detach().clone().requires_grad_(True)
    ref = _attn_ref(qf, kf, vf, heads, scale, causal, key_mask)
    valid = torch.ones(b, sq, dtype=torch.bool, device="cuda")
    if masked and causal:
        valid[0, :17] = False  # fully masked query rows: undefined in the reference too
    _close(o[valid], ref[valid], atol=0.02, rtol=0.02, what="attn fwd")
    d_o = _rand(b, sq, hd, seed=23)
    d_o = d_o * valid[:, :, None]
    torch.nan_to_num(ref, nan=0.0).backward(d_o.float())
    dq, dk, dv = ops.attention_bwd(q, k, v, o, lse, d_o, heads, scale, causal=causal, key_mask=key_mask)
    # head dims that are a multiple of 16 run the tcgen05 / TMEM backward (attention_bwd_tcgen05.cu)
    assert ops.attention_bwd_uses_tcgen05(q, k, v, o, lse, d_o, heads, scale, causal=causal,
                                          key_mask=key_mask) == (d % 16 == 0 and os.environ.get("VB_ATTN_BWD_TC") != "0")
    for got, want, nm in ((dq, qf.grad, "dq"), (dk, kf.grad, "dk"), (dv, vf.grad, "dv")):
        want = torch.nan_to_num(want, nan=0.0)
        if nm == "dq":
            got, want = got[valid], want[valid]
        _close(got, want, atol=0.03 + 0.02 * want.abs().max().item(), rtol=0.03, what=f"attn {nm}")


@pytest.mark.parametrize("b,heads,d,s", [(3, 16, 88, 257), (2, 4, 64, 128), (2, 2, 128, 272),
                                          (1, 3, 40, 200), (19, 16, 88, 257), (2, 5, 96, 65)])
def test_attention_tcgen05_vit_class(b, heads, d, s):
    """Non-causal, unmasked, no-lse self attention with S <= 272 takes the TMEM kernel."""
    ops = _ops()
    hd = heads * d
    qkv = _rand(b, s, 3 * hd, seed=70 + d)
    q, k, v = qkv[:, :, :hd], qkv[:, :, hd:2 * hd], qkv[:, :, 2 * hd:]
    assert ops.attention_uses_tcgen05(q, k, v, heads)
    assert not ops.attention_uses_tcgen05(q, k, v, heads, need_lse=True)
    scale = d ** -0.5
    o = ops.attention(q, k, v, heads, scale)
    ref = _attn_ref(q, k, v, heads, scale, False, None)
    _close(o, ref, atol=0.02, rtol=0.02, what="attn tcgen05")
    o2, _ = ops.attention(q, k, v, heads, scale, need_lse=True)  # the flash kernel (tcgen05 or mma.sync), same inputs
    _close(o, o2, atol=0.02, rtol=0.02, what="single-pass vs flash kernel")


@pytest.mark.parametrize("b,heads,d,sq,skv,causal", [(3, 4, 88, 257, 257, False), (2, 2, 64, 40, 300, False),
                                                     (2, 3, 80, 64, 64, True)])
def test_attention_probs_maps(b, heads, d, sq, skv, causal):
    """vb_attention_probs: the `attentions` output (eilev/model/v2.py:87-95) = softmax(scale q k^T + masks)."""
    ops = _ops()
    hd = heads * d
    q, k = _rand(b, sq, hd, seed=90), _rand(b, skv, hd, seed=91)
    key_mask = torch.ones(b, skv, dtype=torch.uint8, device="cuda")
    key_mask[0, skv - 5:] = 0
    scale = d ** -0.5
    probs = ops.attention_probs(q, k, heads, scale, causal=causal, key_mask=key_mask)
    qf = q.float().view(b, sq, heads, d).transpose(1, 2)
    kf = k.float().view(b, skv, heads, d).transpose(1, 2)
    logits = qf @ kf.transpose(-1, -2) * scale
    logits = logits.masked_fill(key_mask[:, None, None, :] == 0, float("-inf"))
    if causal:
        logits = logits.masked_fill(~torch.ones(sq, skv, dtype=torch.bool, device="cuda").tril(skv - sq), float("-inf"))
    ref = torch.softmax(logits, dim=-1)
    assert probs.shape == (b, heads, sq, skv) and probs.dtype == torch.float32
    _close(probs, ref, atol=2e-5, rtol=1e-4, what="attention maps")
    assert torch.allclose(probs.sum(-1), torch.ones_like(probs.sum(-1)), atol=1e-4)


def test_patch_gather_and_cls():
    ops = _ops()
    px = torch.randn(2, 3, 4, 28, 28, device="cuda")
    out = ops.patch_gather(px, 14, 592)
    frames = px.permute(0, 2, 1, 3, 4).flatten(end_dim=1)
    ref = torch.nn.functional.unfold(frames, kernel_size=14, stride=14).transpose(1, 2).reshape(-1, 588)
    _close(out[:, :588], ref, atol=0.02, rtol=0.01, what="patch gather")
    assert out[:, 588:].abs().max().item() == 0.0
    # patch 12 on 28x28: the strided conv drops the remainder
    out12 = ops.patch_gather(px.to(torch.bfloat16), 12, 432)
    ref12 = torch.nn.functional.unfold(frames[:, :, :24, :24], kernel_size=12, stride=12).transpose(1, 2).reshape(-1, 432)
    _close(out12, ref12.to(torch.bfloat16), atol=1e-6, rtol=0, what="patch gather 12")
    hidden = torch.zeros(5, 7, 16, dtype=torch.bfloat16, device="cuda")
    cls, pos = _rand(16, seed=30), _rand(7, 16, seed=31)
    ops.cls_rows(cls, pos, hidden)
    _close(hidden[:, 0], (cls.float() + pos[0].float()).expand(5, 16), atol=0.02, rtol=0.01)


def test_embed_splice_and_bwd():
    ops = _ops()
    torch.manual_seed(0)
    b, l, dim, vocab, nq = 2, 19, 24, 50, 4
    ids = torch.randint(0, vocab, (b, l), device="cuda")
    attn = torch.ones(b, l, dtype=torch.int64, device="cuda")
    attn[0, :3] = 0
    vmask = torch.zeros(b, l, dtype=torch.int64, device="cuda")
    vmask[0, 4:8] = 1
    vmask[1, 1:5] = 1
    vmask[1, 9:13] = 1
    emb = _rand(vocab, dim, seed=40)
    feats = _rand(3 * nq, dim, seed=41)
    ptab = _rand(l + 2, dim, seed=42)
    e, h, slot, pos, status = ops.embed_splice(ids, attn, vmask, emb, feats, ptab, 2)
    ref_e = emb[ids].clone()
    ref_e[vmask.bool()] = feats
    assert torch.equal(e, ref_e)
    ref_pos = (torch.cumsum(attn, 1) * attn - 1) + 2
    assert torch.equal(pos.view(b, l).long(), ref_pos)
    _close(h, ref_e.float() + ptab[ref_pos].float(), atol=0.02, rtol=0.01)
    assert status.tolist() == [0, 12]
    d_e = _rand(b, l, dim, seed=43)
    d_f = ops.splice_bwd(d_e, slot, 3 * nq)
    assert torch.equal(d_f, d_e[vmask.bool()])
    _, _, _, _, st2 = ops.embed_splice(ids, attn, vmask, emb, feats[:8], ptab, 2)
    assert st2.tolist()[0] == 1


@pytest.mark.parametrize("b,l", [(1, 976), (3, 1500), (5, 2049), (2, 7)])
def test_embed_splice_indices_at_sequence_lengths_beyond_one_chunk(b, l):
    """slot indices (rank among the masked positions, row-major) and OPT positions (cumsum(mask) * mask - 1 + 2)
    from the block-wide scans, on random masks with left padding."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(l)
    dim, vocab = 16, 40
    ids = torch.randint(0, vocab, (b, l), device="cuda", generator=g)
    attn = (torch.rand(b, l, device="cuda", generator=g) > 0.1).long()
    attn[0, : l // 7] = 0
    vmask = ((torch.rand(b, l, device="cuda", generator=g) > 0.6) & (attn == 1)).long()
    nfeat = int(vmask.sum())
    emb, feats, ptab = _rand(vocab, dim, seed=1), _rand(max(nfeat, 1), dim, seed=2), _rand(l + 2, dim, seed=3)
    e, h, slot, pos, status = ops.embed_splice(ids, attn, vmask, emb, feats, ptab, 2)
    assert status.tolist() == [0, nfeat]
    want_slot = torch.where(vmask.bool().flatten(), torch.cumsum(vmask.flatten(), 0) - 1, torch.full((b * l,), -1, device="cuda"))
    assert torch.equal(slot.flatten().long(), want_slot)
    assert torch.equal(pos.view(b, l).long(), torch.cumsum(attn, 1) * attn - 1 + 2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_cross_entropy(dtype):
    ops = _ops()
    torch.manual_seed(1)
    b, l, v = 2, 13, 1003
    logits = (torch.randn(b, l, v, device="cuda") * 3).to(dtype)
    labels = torch.full((b, l), -100, dtype=torch.int64, device="cuda")
    labels[0, 5:9] = torch.randint(0, v, (4,), device="cuda")
    labels[1, 10:] = torch.randint(0, v, (3,), device="cuda")
    loss, row_lse, n_valid = ops.cross_entropy(logits, labels)
    lf = logits.float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lf[:, :-1].reshape(-1, v), labels[:, 1:].reshape(-1), ignore_index=-100)
    assert n_valid.item() == 7
    assert abs(loss.item() - ref.item()) < 2e-3, (loss.item(), ref.item())
    ref.backward()
    gs = torch.full((), 0.5, device="cuda")
    dl = ops.cross_entropy_bwd(logits, labels, row_lse, n_valid, gs)
    _close(dl.reshape(b, l, v), 0.5 * lf.grad, atol=2e-3, rtol=0.02, what="ce bwd")


def test_transpose_convert_misc():
    ops = _ops()
    x = _rand(100, 37, seed=50)
    assert torch.equal(ops.transpose(x), x.t())
    xs = _rand(64, 96, seed=51)[:, :40]
    assert torch.equal(ops.transpose(xs), xs.t())
    f = torch.randn(1000, device="cuda")
    assert torch.equal(ops.convert(f, torch.bfloat16), f.to(torch.bfloat16))
    assert torch.equal(ops.convert(f.to(torch.bfloat16), torch.float32), f.to(torch.bfloat16).float())
    dy, pre = _rand(50, 33, seed=52), _rand(50, 33, seed=53)
    pf = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(pf).backward(dy.float())
    _close(ops.act_bwd(dy, pre, ops.EPI_GELU), pf.grad, atol=0.02, rtol=0.02)
    post = torch.relu(pre)
    _close(ops.act_bwd(dy, post, ops.EPI_RELU), dy.float() * (post > 0), atol=1e-6, rtol=0)
    big = _rand(5000, 70, seed=54)
    _close(ops.colsum(big), big.float().sum(0), atol=0.5, rtol=0.01)
    _close(ops.add(dy, pre), dy.float() + pre.float(), atol=0.02, rtol=0.01)


def test_dropout_mask_is_shared_by_gemm_epilogue_and_elementwise_kernel():
    ops = _ops()
    seed = torch.tensor([12345], dtype=torch.int64, device="cuda")
    p, salt = 0.1, 77
    m, n, k = 300, 512, 256
    ones = torch.ones(m, n, dtype=torch.bfloat16, device="cuda")
    mask = ops.dropout(ones, p, seed, salt).float()          # 0 or 1/(1-p)
    keep = (mask > 0).float().mean().item()
    assert abs(keep - 0.9) < 0.01, keep
    assert torch.allclose(mask[mask > 0], torch.tensor(1 / 0.9, device="cuda"), atol=5e-3)
    assert not torch.equal(mask, ops.dropout(ones, p, seed, salt + 1).float())
    a, w = _rand(m, k, scale=0.5, seed=80), _rand(n, k, scale=0.1, seed=81)
    bias = torch.randn(n, device="cuda")
    res = _rand(m, n, seed=82)
    ref = torch.relu(a.float() @ w.float().t() + bias) * mask + res.float()
    for be in (ops.GEMM_TCGEN05, ops.GEMM_GENERIC):
        out = ops.gemm(a, w, bias, residual=res, epilogue=ops.EPI_RELU, backend=be, dropout=(p, seed, salt))
        _close(out, ref, atol=0.03, rtol=0.01, what=f"gemm dropout backend {be}")
    seed.add_(1)  # a new step draws a new mask
    assert not torch.equal(mask, ops.dropout(ones, p, seed, salt).float())


@pytest.mark.parametrize("b,heads,d,sq,skv,causal", [(2, 12, 64, 32, 32, False), (2, 12, 64, 32, 300, False),
                                                      (1, 4, 80, 200, 200, True)])
def test_attention_dropout_fwd_bwd(b, heads, d, sq, skv, causal):
    ops = _ops()
    hd = heads * d
    scale = d ** -0.5
    seed = torch.tensor([999], dtype=torch.int64, device="cuda")
    p, salt = 0.1, 5
    q = _rand(b, sq, hd, seed=90)
    kv = _rand(b, skv, 2 * hd, seed=91)
    k, v = kv[:, :, :hd], kv[:, :, hd:]
    drop = (p, seed, salt)
    o, lse = ops.attention(q, k, v, heads, scale, causal=causal, need_lse=True, dropout=drop)
    mask = ops.dropout(torch.ones(b * heads * sq, skv, dtype=torch.bfloat16, device="cuda"), p, seed, salt)
    mask = mask.float().view(b, heads, sq, skv)
    qf = q.float().detach().clone().requires_grad_(True)
    kf = k.float().detach().clone().requires_grad_(True)
    vf = v.float().detach().clone().requires_grad_(True)
    qh = qf.view(b, sq, heads, d).transpose(1, 2)
    kh = kf.view(b, skv, heads, d).transpose(1, 2)
    vh = vf.view(b, skv, heads, d).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * scale
    if causal:
        i = torch.arange(sq, device="cuda")[:, None]
        j = torch.arange(skv, device="cuda")[None, :]
        s = s.masked_fill(j > i + (skv - sq), -1e30)
    ref = ((torch.softmax(s, -1) * mask) @ vh).transpose(1, 2).reshape(b, sq, hd)
    _close(o, ref, atol=0.03, rtol=0.02, what="attn dropout fwd")
    d_o = _rand(b, sq, hd, seed=92)
    ref.backward(d_o.float())
    dq, dk, dv = ops.attention_bwd(q, k, v, o, lse, d_o, heads, scale, causal=causal, dropout=drop)
    for got, want, nm in ((dq, qf.grad, "dq"), (dk, kf.grad, "dk"), (dv, vf.grad, "dv")):
        _close(got, want, atol=0.03 + 0.02 * want.abs().max().item(), rtol=0.03, what=f"attn dropout {nm}")


def test_adamw_matches_torch():
    ops = _ops()
    torch.manual_seed(2)
    p = torch.randn(10000, device="cuda")
    ref_p = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref_p], lr=1e-2, weight_decay=0.05)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn(10000, device="cuda")
        ref_p.grad = g.clone()
        opt.step()
        ops.adamw_(p, g, m, v, lr=1e-2, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.05, step=step)
    _close(p, ref_p.detach(), atol=1e-5, rtol=1e-5, what="adamw")
    acc = torch.zeros((), device="cuda")
    ops.sumsq(p, acc)
    assert abs(acc.item() - (p.double() ** 2).sum().item()) < 1e-2 * acc.item()
    # the global-norm reduction is deterministic (data-parallel replicas must derive the same clip coefficient):
    # a large buffer, repeated launches, accumulation into a non-zero `out`
    big = torch.randn(3_000_001, device="cuda")
    first = None
    for _ in range(20):
        acc = torch.zeros((), device="cuda")
        ops.sumsq(big, acc)
        first = acc.clone() if first is None else first
        assert torch.equal(acc, first), (acc.item(), first.item())
    assert abs(first.item() - (big.double() ** 2).sum().item()) < 1e-5 * first.item()
    acc = torch.full((), 5.0, device="cuda")
    ops.sumsq(p, acc)
    assert abs(acc.item() - 5.0 - (p.double() ** 2).sum().item()) < 1e-4 * acc.item()


def test_gemv_and_paged_decode():
    ops = _ops()
    for m in (1, 3, 8):
        x, w = _rand(m, 2560, seed=60), _rand(7680, 2560, scale=0.05, seed=61)
        bias = torch.randn(7680, device="cuda")
        y = ops.gemv(x, w, bias, alpha=0.25, alpha_cols=2560)
        ref = x.float() @ w.float().t() + bias
        ref[:, :2560] *= 0.25
        _close(y, ref, atol=0.05, rtol=0.02, what="gemv")
        g, bt = torch.randn(2560, device="cuda"), torch.randn(2560, device="cuda")
        res = _rand(m, 7680, seed=64)
        y2 = ops.gemv(x, w, bias, residual=res, epilogue=ops.EPI_RELU, ln=(g, bt, 1e-5))
        xn = torch.nn.functional.layer_norm(x.float(), (2560,), g, bt, 1e-5).to(torch.bfloat16).float()
        _close(y2, torch.relu(xn @ w.float().t() + bias) + res.float(), atol=0.08, rtol=0.02, what="ln+gemv")
    xo, wo = _rand(2, 200, seed=65), _rand(33, 200, seed=66)  # odd shapes: legacy kernel
    _close(ops.gemv(xo, wo), xo.float() @ wo.float().t(), atol=0.05, rtol=0.02, what="gemv odd")
    heads, d, page, b, l = 4, 80, 16, 2, 37
    hd = heads * d
    max_pages = 4
    kc = torch.zeros(b * max_pages, page, hd, dtype=torch.bfloat16, device="cuda")
    vc = torch.zeros_like(kc)
    table = torch.arange(b * max_pages, dtype=torch.int32, device="cuda").view(b, max_pages).flip(0).contiguous()
    kv = _rand(b, l, 2 * hd, seed=62)
    ops.paged_kv_write(kv[:, :, :hd], kv[:, :, hd:], kc, vc, table, page)
    qkv = _rand(b, 3 * hd, seed=63)
    ctx = torch.full((b,), l + 1, dtype=torch.int32, device="cuda")
    first = torch.tensor([5, 0], dtype=torch.int32, device="cuda")
    out = ops.paged_decode_attention(qkv, kc, vc, table, ctx, first, heads, page, 1.0)
    k_all = torch.cat([kv[:, :, :hd], qkv[:, None, hd:2 * hd]], 1)
    v_all = torch.cat([kv[:, :, hd:], qkv[:, None, 2 * hd:]], 1)
    mask = torch.ones(b, l + 1, dtype=torch.uint8, device="cuda")
    mask[0, :5] = 0
    ref = _attn_ref(qkv[:, None, :hd], k_all, v_all, heads, 1.0, False, mask)[:, 0]
    _close(out, ref, atol=0.02, rtol=0.02, what="paged decode")


@pytest.mark.parametrize("m,n,k", [(1, 2560, 10240), (1, 1003, 2560), (2, 10240, 2560), (4, 50272, 2560),
                                   (8, 2560, 2560), (16, 2560, 2560), (16, 3000, 1024), (1, 77, 256),
                                   (3, 2560, 10240), (8, 640, 10240), (16, 2560, 10240), (11, 1003, 10240)])
def test_gemv_bulk_ring_shapes(m, n, k):
    """Weight-streaming GEMV (bulk-copy ring kernel and its fall-backs) across the decode
    shapes: ragged N, every M bucket, K walked in slices, f32 output, residual + bias."""
    ops = _ops()
    x, w = _rand(m, k, seed=70), _rand(n, k, scale=0.03, seed=71)
    bias = torch.randn(n, device="cuda")
    res = _rand(m, n, seed=72)
    ref = x.float() @ w.float().t() + bias
    tol = dict(atol=0.02 * (k / 2560) ** 0.5 + 0.03, rtol=0.02)
    _close(ops.gemv(x, w, bias), ref, what="gemv", **tol)
    _close(ops.gemv(x, w, bias, residual=res, epilogue=ops.EPI_RELU), torch.relu(ref) + res.float(),
           what="gemv relu+res", **tol)
    y32 = ops.gemv(x, w, out_dtype=torch.float32)
    assert y32.dtype == torch.float32
    _close(y32, x.float() @ w.float().t(), atol=2e-3 * (k / 256) ** 0.5, rtol=1e-3, what="gemv f32")
    if k <= 4096:
        g, bt = torch.randn(k, device="cuda"), torch.randn(k, device="cuda")
        xn = torch.nn.functional.layer_norm(x.float(), (k,), g, bt, 1e-5).to(torch.bfloat16).float()
        _close(ops.gemv(x, w, bias, ln=(g, bt, 1e-5)), xn @ w.float().t() + bias, what="ln+gemv",
               atol=0.08, rtol=0.02)


def test_gemv_back_to_back_pdl_chain():
    """A chain of dependent GEMVs on one stream (programmatic dependent launch lets each
    prefetch weights under its predecessor): results must equal the serial composition."""
    ops = _ops()
    x = _rand(1, 2560, seed=80)
    ws = [_rand(2560, 2560, scale=0.02, seed=81 + i) for i in range(6)]
    y = x
    ref = x.float()
    for w in ws:
        y = ops.gemv(y, w, residual=y)
        ref = (ref @ w.float().t() + ref).to(torch.bfloat16).float()
    _close(y, ref, atol=0.05, rtol=0.03, what="gemv chain")


def test_decode_embed_advances_counters():
    ops = _ops()
    emb, pos = _rand(100, 256, seed=90), _rand(40, 256, seed=91)
    tok = torch.tensor([3, 99, 0], device="cuda")
    nv = torch.tensor([5, 0, 30], dtype=torch.int32, device="cuda")
    cl = torch.tensor([9, 1, 31], dtype=torch.int32, device="cuda")
    x = ops.decode_embed(tok, emb, pos, nv, cl, 2)
    ref = emb[tok].float() + pos[torch.tensor([7, 2, 32], device="cuda")].float()
    _close(x, ref, atol=0.02, rtol=0.01, what="decode_embed")
    assert nv.tolist() == [6, 1, 31] and cl.tolist() == [10, 2, 32]


def test_paged_decode_long_context_splits():
    ops = _ops()
    heads, d, page, b, l = 8, 80, 64, 2, 1000
    hd = heads * d
    max_pages = 17
    kc = torch.zeros(b * max_pages, page, hd, dtype=torch.bfloat16, device="cuda")
    vc = torch.zeros_like(kc)
    table = torch.arange(b * max_pages, dtype=torch.int32, device="cuda").view(b, max_pages).contiguous()
    kv = _rand(b, l, 2 * hd, seed=92)
    ops.paged_kv_write(kv[:, :, :hd], kv[:, :, hd:], kc, vc, table, page)
    qkv = _rand(b, 3 * hd, seed=93)
    ctx = torch.full((b,), l + 1, dtype=torch.int32, device="cuda")
    first = torch.tensor([17, 0], dtype=torch.int32, device="cuda")
    k_all = torch.cat([kv[:, :, :hd], qkv[:, None, hd:2 * hd]], 1)
    v_all = torch.cat([kv[:, :, hd:], qkv[:, None, 2 * hd:]], 1)
    mask = torch.ones(b, l + 1, dtype=torch.uint8, device="cuda")
    mask[0, :17] = 0
    ref = _attn_ref(qkv[:, None, :hd], k_all, v_all, heads, 1.0, False, mask)[:, 0]
    for splits in (1, 8, 16, 33):
        out = ops.paged_decode_attention(qkv, kc, vc, table, ctx, first, heads, page, 1.0, splits=splits)
        _close(out, ref, atol=0.02, rtol=0.02, what=f"paged decode splits={splits}")


def test_attention_merge_equals_joint_softmax():
    """classify(): attention over [prompt keys ; own continuation] == merge of the two partials."""
    ops = _ops()
    heads, d, b, ncls, lc, lp = 4, 80, 2, 3, 5, 37
    hd = heads * d
    q = _rand(b * ncls, lc, hd, seed=100)
    kc_, vc_ = _rand(b * ncls, lc, hd, seed=101), _rand(b * ncls, lc, hd, seed=102)
    kp, vp = _rand(b, lp, hd, seed=103), _rand(b, lp, hd, seed=104)
    pmask = torch.ones(b, lp, dtype=torch.uint8, device="cuda")
    pmask[1, :9] = 0
    o1, l1 = ops.attention(q.view(b, ncls * lc, hd), kp, vp, heads, 0.3, key_mask=pmask, need_lse=True)
    o2, l2 = ops.attention(q, kc_, vc_, heads, 0.3, causal=True, need_lse=True)
    got = ops.attention_merge(o1, l1, o2, l2, heads)
    k_all = torch.cat([kp.repeat_interleave(ncls, 0), kc_], 1)
    v_all = torch.cat([vp.repeat_interleave(ncls, 0), vc_], 1)
    m_all = torch.cat([pmask.repeat_interleave(ncls, 0), torch.ones(b * ncls, lc, dtype=torch.uint8, device="cuda")], 1)
    ref = _attn_ref(q, k_all, v_all, heads, 0.3, True, m_all)
    _close(got, ref, atol=0.02, rtol=0.02, what="attention merge")


def test_token_logprob():
    ops = _ops()
    logits = torch.randn(7, 1000, device="cuda") * 3
    tg = torch.tensor([0, 999, -100, 5, 1000, 17, 3], device="cuda")
    got = ops.token_logprob(logits, tg)
    ref = torch.log_softmax(logits, -1)
    for i, t in enumerate(tg.tolist()):
        want = ref[i, t].item() if 0 <= t < 1000 else 0.0
        assert abs(got[i].item() - want) < 1e-4, (i, got[i].item(), want)
    rows = torch.tensor([6, 6, 0], device="cuda")
    got2 = ops.token_logprob(logits.to(torch.bfloat16), torch.tensor([1, 2, 3], device="cuda"), rows)
    ref2 = torch.log_softmax(logits.to(torch.bfloat16).float(), -1)
    assert torch.allclose(got2, torch.stack([ref2[6, 1], ref2[6, 2], ref2[0, 3]]), atol=1e-3)


def test_rmsnorm_fwd_bwd_matches_torch():
    ops = _ops()
    for rows, cols in ((37, 2048), (5, 128), (9, 72)):
        x = _rand(rows, cols, seed=110)
        g = (1 + 0.1 * torch.randn(cols, device="cuda")).float()
        y, rstd = ops.rmsnorm(x, g, 1e-6, save_stats=True)
        xf = x.float().requires_grad_(True)
        ref = g * xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)
        _close(y, ref, atol=0.03, rtol=0.02, what="rmsnorm")
        dy = _rand(rows, cols, seed=111)
        add = _rand(rows, cols, seed=112)
        ref.backward(dy.float())
        dx = ops.rmsnorm_bwd(dy, x, g, rstd, dx_add=add)
        _close(dx, xf.grad + add.float(), atol=0.04, rtol=0.03, what="rmsnorm bwd")


def test_gated_gelu_fwd_bwd_matches_torch():
    ops = _ops()
    rows, dff = 33, 256
    h01 = _rand(rows, 2 * dff, seed=113) * 2
    out = ops.gated_gelu(h01)
    hf = h01.float().requires_grad_(True)
    ref = torch.nn.functional.gelu(hf[:, :dff], approximate="tanh") * hf[:, dff:]
    _close(out, ref, atol=0.03, rtol=0.02, what="gated gelu")
    d_out = _rand(rows, dff, seed=114)
    ref.backward(d_out.float())
    _close(ops.gated_gelu_bwd(d_out, h01), hf.grad, atol=0.04, rtol=0.03, what="gated gelu bwd")


@pytest.mark.parametrize("causal,sq,skv", [(False, 104, 104), (True, 9, 9), (False, 9, 104), (True, 70, 70)])
def test_attention_relative_bias_fwd_bwd(causal, sq, skv):
    """T5 attention: unscaled scores + per-head relative-position table (+ masks)."""
    ops = _ops()
    b, heads, d = 2, 3, 64
    hd = heads * d
    q, k, v = _rand(b, sq, hd, seed=120), _rand(b, skv, hd, seed=121), _rand(b, skv, hd, seed=122)
    tab = torch.randn(heads, sq + skv - 1, device="cuda")
    mask = torch.ones(b, skv, dtype=torch.uint8, device="cuda")
    if not causal:
        mask[1, skv - 5:] = 0
    km = None if causal else mask
    o, lse = ops.attention(q, k, v, heads, 0.2, causal=causal, key_mask=km, need_lse=True, rel_bias=tab)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    qh = qf.view(b, sq, heads, d).transpose(1, 2)
    kh = kf.view(b, skv, heads, d).transpose(1, 2)
    vh = vf.view(b, skv, heads, d).transpose(1, 2)
    i = torch.arange(sq, device="cuda")[:, None]
    j = torch.arange(skv, device="cuda")[None, :]
    s = qh @ kh.transpose(-1, -2) * 0.2 + tab[:, (j - i) + (sq - 1)][None]
    if causal:
        s = s.masked_fill(j > i + (skv - sq), -1e30)
    else:
        s = s.masked_fill(mask[:, None, None, :] == 0, -1e30)
    ref = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(b, sq, hd)
    _close(o, ref, atol=0.03, rtol=0.03, what="attention rel_bias")
    d_o = _rand(b, sq, hd, seed=123)
    ref.backward(d_o.float())
    dq, dk, dv = ops.attention_bwd(q, k, v, o, lse, d_o, heads, 0.2, causal=causal, key_mask=km, rel_bias=tab)
    _close(dq, qf.grad, atol=0.05, rtol=0.05, what="rel_bias dq")
    _close(dk, kf.grad, atol=0.05, rtol=0.05, what="rel_bias dk")
    _close(dv, vf.grad, atol=0.05, rtol=0.05, what="rel_bias dv")


def test_cross_entropy_unshifted():
    ops = _ops()
    b, l, v = 2, 9, 264
    logits = (torch.randn(b, l, v, device="cuda") * 2).to(torch.bfloat16)
    labels = torch.randint(0, v, (b, l), device="cuda")
    labels[1, 6:] = -100
    loss, row_lse, n_valid = ops.cross_entropy(logits, labels, shift=0)
    lf = logits.float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lf.view(-1, v), labels.view(-1), ignore_index=-100)
    assert abs(float(loss) - float(ref)) < 2e-3 and int(n_valid) == 15
    ref.backward()
    d = ops.cross_entropy_bwd(logits, labels, row_lse, n_valid, None, shift=0)
    _close(d.view(b, l, v), lf.grad, atol=2e-3, rtol=0.05, what="ce bwd unshifted")


def test_embedding_gather():
    ops = _ops()
    table = _rand(50, 128, seed=130)
    ids = torch.tensor([[0, 49, 7], [3, 3, 60]], device="cuda")
    out = ops.embedding(ids, table)
    assert torch.equal(out, table[ids.clamp(max=49)])


@pytest.mark.parametrize("m,n,k", [(1, 6144, 2048), (3, 384, 128), (8, 2048, 2048), (2, 300, 192)])
def test_gemv_fused_rmsnorm(m, n, k):
    """vb_gemv with ln_gamma and no ln_beta = T5LayerNorm (RMSNorm) fused into the x staging."""
    ops = _ops()
    x, w = _rand(m, k, seed=150), _rand(n, k, scale=0.03, seed=151)
    g = (1 + 0.1 * torch.randn(k, device="cuda")).float()
    xf = x.float()
    xn = (g * xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).to(torch.bfloat16).float()
    got = ops.gemv(x, w, ln=(g, None, 1e-6), out_dtype=torch.float32)
    _close(got, xn @ w.float().t(), atol=0.03, rtol=0.02, what="rms+gemv")


@pytest.mark.parametrize("splits", [1, 4, 8])
def test_decode_cross_attention_over_dense_encoder_kv(splits):
    """One query per sequence over K / V that are column slices of one (B, L, n*H*D) projection
    output, valid tokens [first, end) per row (right- and left-padded rows)."""
    ops = _ops()
    b, heads, d, l, n_slices = 3, 4, 64, 300, 6
    hd = heads * d
    ckv = _rand(b, l, n_slices * hd, seed=160)
    k, v = ckv[:, :, 2 * hd:3 * hd], ckv[:, :, 3 * hd:4 * hd]
    q = _rand(b, hd, seed=161)
    first = torch.tensor([0, 17, 0], dtype=torch.int32, device="cuda")
    end = torch.tensor([l, l, 211], dtype=torch.int32, device="cuda")
    mask = torch.zeros(b, l, dtype=torch.uint8, device="cuda")
    for i in range(b):
        mask[i, int(first[i]):int(end[i])] = 1
    ws = torch.empty(b * heads * splits * (d + 2), dtype=torch.float32, device="cuda")
    cnt = torch.zeros(b * heads, dtype=torch.int32, device="cuda")
    seq = torch.arange(b, dtype=torch.int32, device="cuda")
    for _ in range(2):  # twice: the counters must come back to zero
        out = ops.decode_cross_attention(q, k, v, seq, end, first, heads, 0.7, workspace=ws, counters=cnt, splits=splits)
    ref = _attn_ref(q[:, None, :], k, v, heads, 0.7, False, mask)[:, 0]
    _close(out, ref, atol=0.02, rtol=0.02, what="decode cross attention")
    assert int(cnt.abs().sum()) == 0


@pytest.mark.parametrize("rows,cols", [(976, 2560), (37, 80), (5, 24)])
def test_layernorm_backward_with_fused_dropout_output(rows, cols):
    """vb_layernorm_bwd_dropout: dx as vb_layernorm_bwd gives it, and dropout(dx) as vb_dropout gives it."""
    ops = _ops()
    dy, x, add = _rand(rows, cols, seed=61), _rand(rows, cols, seed=62), _rand(rows, cols, seed=63)
    gamma = 1.0 + 0.2 * torch.randn(cols, device="cuda")
    beta = torch.zeros(cols, device="cuda")
    _, mean, rstd = ops.layernorm(x, gamma, beta, 1e-5, save_stats=True)
    seed = torch.tensor([12345], dtype=torch.int64, device="cuda")
    dx_ref = ops.layernorm_bwd(dy, x, gamma, mean, rstd, dx_add=add)
    dx, dxm = ops.layernorm_bwd(dy, x, gamma, mean, rstd, dx_add=add, dropout=(0.1, seed, 9))
    assert torch.equal(dx, dx_ref)
    assert torch.equal(dxm, ops.dropout(dx_ref, 0.1, seed, 9))


def test_attention_dropout_and_bias_on_the_other_kernels_in_a_subprocess():
    """Dropout on the probabilities and the T5 relative bias run on the tcgen05 flash kernels by default (the tests
    above); VB_ATTN_TC_SLOW=0 routes them to the mma.sync kernels, which stay covered by re-running the same parity
    tests with the switch (it is read once per process)."""
    import subprocess
    import sys
    if os.environ.get("VB_ATTN_TC_SLOW") == "0":
        pytest.skip("already the forced run")
    env = dict(os.environ, VB_ATTN_TC_SLOW="0")
    r = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-x", "-k",
                        "attention_dropout_fwd_bwd or attention_relative_bias_fwd_bwd"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]

