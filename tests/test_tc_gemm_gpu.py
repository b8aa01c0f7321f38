"""tcgen05 GEMM (csrc/tc_gemm.cu) against torch fp32 matmul of the same bf16 operands: all four operand layouts, both tile
widths, ragged edges (TMA zero-fill / clipped stores), split-K wgrad accumulation, every epilogue flag."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(M, N, K, tA, tB, seed=0):
    torch.manual_seed(seed)
    A = torch.randn((K, M) if tA else (M, K), device='cuda').to(torch.bfloat16)
    B = torch.randn((N, K) if tB else (K, N), device='cuda').to(torch.bfloat16)
    ref = (A.float().t() if tA else A.float()) @ (B.float().t() if tB else B.float())
    return A, B, ref


@pytest.mark.parametrize('tA,tB', [(0, 1), (0, 0), (1, 0), (1, 1)])
@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (128, 256, 128), (256, 512, 512), (384, 128, 1024), (200, 136, 72), (64, 1192, 512),
                                   (1000, 520, 200), (130, 264, 16)])
def test_tc_gemm_layouts(ops, tA, tB, M, N, K):
    if (tA and M % 8) or (not tA and K % 8) or (tB and K % 8) or (not tB and N % 8):
        pytest.skip('leading dimension not 16-byte aligned: SIMT path')
    A, B, ref = _mk(M, N, K, tA, tB)
    for dt, tol in ((torch.float32, 1e-3), (torch.bfloat16, 1e-2)):
        out = ops.gemm(A, B, transA=bool(tA), transB=bool(tB), out_dtype=dt)
        err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
        assert err < tol, (dt, err)


def test_tc_gemm_is_the_tensor_core_path(ops):
    """The same call with TXL_DISABLE_TC unset must not be bit-identical to exact fp32 FMA order... but must agree closely;
    and big-K accumulation stays in fp32 (no bf16 partial sums)."""
    A, B, ref = _mk(256, 256, 8192, 0, 1, seed=3)
    out = ops.gemm(A, B, transB=True, out_dtype=torch.float32)
    assert (out - ref).abs().max().item() / ref.abs().max().item() < 2e-5


def test_tc_gemm_splitk_accumulate(ops):
    """wgrad shape: dW[n_out, n_in] += dY^T X with the token dimension as K — few tiles, so split-K with fp32 reductions."""
    torch.manual_seed(1)
    tokens, n_out, n_in = 4096, 384, 256
    dY = torch.randn(tokens, n_out, device='cuda').to(torch.bfloat16)
    X = torch.randn(tokens, n_in, device='cuda').to(torch.bfloat16)
    G = torch.ones(n_out, n_in, device='cuda')
    ops.gemm(dY, X, transA=True, out=G, accumulate=True)
    ref = 1 + dY.float().t() @ X.float()
    assert (G - ref).abs().max().item() / ref.abs().max().item() < 1e-4
    # into a row-slice of a larger gradient (the qkv_net k/v rows)
    big = torch.zeros(3 * n_out, n_in, device='cuda')
    ops.gemm(dY, X, transA=True, out=big[n_out:2 * n_out], accumulate=True)
    assert (big[n_out:2 * n_out] - (ref - 1)).abs().max().item() / ref.abs().max().item() < 1e-4
    assert big[:n_out].abs().sum() == 0 and big[2 * n_out:].abs().sum() == 0


def test_tc_gemm_epilogues(ops):
    torch.manual_seed(2)
    M, N, K = 300, 512, 256
    A, B, ref0 = _mk(M, N, K, 0, 1, seed=2)
    bias = torch.randn(N, device='cuda')
    ref = torch.relu(ref0 + bias)
    cs = torch.zeros(N, device='cuda')
    out = ops.gemm(A, B, transB=True, bias=bias, relu=True, colsum=cs)
    assert out.dtype == torch.bfloat16
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(cs, ref.sum(0), rtol=1e-3, atol=0.5)
    acc = torch.ones(M, N, device='cuda', dtype=torch.bfloat16)
    ops.gemm(A, B, transB=True, out=acc, accumulate=True)
    torch.testing.assert_close(acc.float(), 1 + ref0, rtol=1e-2, atol=2e-2)
    acc32 = torch.ones(M, N, device='cuda')
    ops.gemm(A, B, transB=True, out=acc32, accumulate=True, bias=bias)     # bias => not the split-K path
    torch.testing.assert_close(acc32, 1 + ref0 + bias, rtol=1e-4, atol=1e-3)
    aux = torch.randn(M, N, device='cuda').to(torch.bfloat16)
    out = ops.gemm(A, B, transB=True, mask_pos_aux=aux)
    torch.testing.assert_close(out.float(), ref0 * (aux.float() > 0), rtol=1e-2, atol=1e-2)
    # backward of relu+dropout from the post-dropout activation: its zeros are the mask, only the 1/(1-p) scale is applied
    out = ops.gemm(A, B, transB=True, mask_pos_aux=aux, drop_p=0.25, seed=5, site=3, aux_is_dropped=True)
    torch.testing.assert_close(out.float(), ref0 * (aux.float() > 0) / 0.75, rtol=1e-2, atol=2e-2)
    # the same mask at one bit per element: written by a forward GEMM (relu + dropout), read by the backward GEMM
    bits = torch.empty((N + 31) // 32, M, dtype=torch.int32, device='cuda')
    hfw = ops.gemm(A, B, transB=True, bias=bias, relu=True, drop_p=0.25, seed=5, site=3, emit_live_bits=bits)
    words = bits.cpu().numpy().astype('uint32')                       # [N/32, M]
    live = ((words[:, :, None] >> torch.arange(32).numpy()[None, None, :]) & 1).transpose(1, 0, 2).reshape(M, -1)[:, :N]
    assert (torch.from_numpy(live.astype('bool')).cuda() == (hfw > 0)).all()
    out2 = ops.gemm(A, B, transB=True, mask_live_bits=bits, drop_p=0.25, seed=5, site=3)
    ref2 = ops.gemm(A, B, transB=True, mask_pos_aux=hfw, drop_p=0.25, seed=5, site=3, aux_is_dropped=True)
    assert torch.equal(out2, ref2)
    d1 = ops.gemm(A, B, transB=True, drop_p=0.25, seed=5, site=3, out_dtype=torch.float32)
    os.environ['TXL_DISABLE_TC'] = '0'
    kept = d1 != 0
    assert 0.72 < kept.float().mean().item() < 0.78
    torch.testing.assert_close(d1[kept], (ref0 / 0.75)[kept], rtol=1e-4, atol=1e-3)
    # strided output view / operand views
    wide = torch.zeros(M, 3 * N, device='cuda', dtype=torch.bfloat16)
    ops.gemm(A, B, transB=True, out=wide[:, N:2 * N])
    torch.testing.assert_close(wide[:, N:2 * N].float(), ref0, rtol=1e-2, atol=1e-2)
    assert wide[:, :N].abs().sum() == 0 and wide[:, 2 * N:].abs().sum() == 0
    out = ops.gemm(wide[:, N:2 * N], B, transA=False, transB=False, out_dtype=torch.float32)      # A column-slice (ld=3N), B as [K,N]
    torch.testing.assert_close(out, wide[:, N:2 * N].float() @ B.float(), rtol=1e-4, atol=1e-2)


def test_tc_gemm_dropout_mask_matches_simt(ops):
    """fwd (tensor-core epilogue) and a SIMT-shaped call must draw the same counter-based mask for the same (seed, site, index)."""
    A, B, _ = _mk(128, 128, 64, 0, 1)
    a = ops.gemm(A, B, transB=True, drop_p=0.5, seed=9, site=1, out_dtype=torch.float32)
    A32, B32 = A.float(), B.float()
    b = ops.gemm(A32, B32, transB=True, drop_p=0.5, seed=9, site=1)
    assert torch.equal(a != 0, b != 0)


def test_tc_gemm_transposed_store_and_row_bias(ops):
    """Decode-time Linears: y^T = W x^T with the sequences as the GEMM N; epilogue stores y (transposed back) and adds the per-feature bias."""
    torch.manual_seed(4)
    for B, N, K in ((64, 1536, 512), (8, 512, 2048), (37, 1190, 512)):
        x = torch.randn(B, K, device='cuda').to(torch.bfloat16)
        W = (0.05 * torch.randn(N, K, device='cuda')).to(torch.bfloat16)
        bias = torch.randn(N, device='cuda')
        ref = torch.relu(x.float() @ W.float().t() + bias)
        Np = (N + 7) // 8 * 8
        buf = torch.zeros(B, Np, device='cuda', dtype=torch.bfloat16)
        y = ops.gemm(W, x, transB=True, bias=bias, relu=True, bias_row=True, transpose_out=True, out=buf[:, :N])
        torch.testing.assert_close(y.float(), ref, rtol=2e-2, atol=2e-2)
        y32 = ops.gemm(W.float(), x.float(), transB=True, bias=bias, relu=True, bias_row=True, transpose_out=True)      # SIMT path, same flags
        torch.testing.assert_close(y32, torch.relu(x.float() @ W.float().t() + bias), rtol=1e-4, atol=1e-4)
