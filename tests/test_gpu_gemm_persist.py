"""GPU tests of the persistent CTA-pair projection kernel (csrc/gemm_persist.cuh) behind gpt_linear_{fwd,dgrad}_tf32[x3]:
against float64 matmuls and, bit for bit, against the one-tile-per-CTA kernel it replaces at large M (same products, same
accumulation order along K: the two kernels must agree exactly).

Tolerances: 3xTF32 <= 1e-5 relative of the largest element (fp32-grade), TF32 <= 2e-3.
"""
import numpy as np
import pytest
import torch

from gcn_over_pruned_trees_b200 import ops, synth

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


@pytest.fixture(autouse=True)
def _restore():
    yield
    ops.gemm_persist_config(2, 65536)


@pytest.mark.parametrize('cg', (1, 2))
@pytest.mark.parametrize('M,N,K', [(70000, 512, 360), (65536, 512, 512), (66000, 200, 360), (65537, 512, 200),
                                   (131072 + 77, 64, 36), (70000, 320, 128)])
def test_persistent_gemm_matches_float64_and_the_small_kernel(M, N, K, cg):
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    x = torch.randn(M, K, device=DEV, generator=g)
    w = torch.randn(N, K, device=DEV, generator=g) / np.sqrt(K)
    dy = torch.randn(M, N, device=DEV, generator=g)
    ref = x.double() @ w.double().t()
    dref = dy.double() @ w.double()
    ws = ops.weight_prep(w, 'tf32x3')
    ops.gemm_persist_config(0, 65536)
    small = {m: (ops.linear_fwd(x, w, m, ws if m == 'tf32x3' else None),
                 ops.linear_dgrad(dy, w, m, ws if m == 'tf32x3' else None)) for m in ('tf32', 'tf32x3')}
    ops.gemm_persist_config(cg, 65536)
    for mode, tol in (('tf32x3', 1e-5), ('tf32', 2e-3)):
        y = ops.linear_fwd(x, w, mode, ws if mode == 'tf32x3' else None)
        dx = ops.linear_dgrad(dy, w, mode, ws if mode == 'tf32x3' else None)
        torch.cuda.synchronize()
        assert _rel(y, ref) <= tol, (mode, 'fwd')
        assert _rel(dx, dref) <= tol, (mode, 'dgrad')
        assert torch.equal(y, small[mode][0]), (mode, 'fwd vs small kernel')
        assert torch.equal(dx, small[mode][1]), (mode, 'dgrad vs small kernel')


@pytest.mark.parametrize('cg', (1, 2))
def test_persistent_dgrad_with_k2_prologue_epilogue(cg):
    """linear_dgrad_masked (K2's backward prologue in the GEMM epilogue) through the persistent kernel == the small one."""
    B, T, H, N2 = 160, 512, 512, 512
    batch = synth.make_batch_torch(3, B, T, device=DEV)
    csr = ops.prune_csr(batch[5], batch[6], batch[7], batch[4], batch[1], -1)
    y = torch.randn(B * T, H, device=DEV)
    rng = torch.tensor([5, 9], dtype=torch.int64, device=DEV)
    _, act = ops.aggregate_fwd(y, csr, torch.zeros(H, device=DEV), drop_p=0.5, rng_state=rng, subseq=0, want_act=True)
    w = torch.randn(N2, H, device=DEV) / np.sqrt(H)
    ws = ops.weight_prep(w, 'tf32x3')
    dnext = torch.randn(B * T, N2, device=DEV)
    ops.gemm_persist_config(0, 65536)
    ref = ops.linear_dgrad_masked(dnext, w, ws, act, csr, 0.5)
    ops.gemm_persist_config(cg, 65536)
    got = ops.linear_dgrad_masked(dnext, w, ws, act, csr, 0.5)
    torch.cuda.synchronize()
    assert torch.equal(got, ref)
    assert float(got.abs().max()) > 0


def test_small_batches_keep_the_one_tile_per_cta_kernel():
    """Below min_rows nothing changes: the TACRED-shape step does not take the persistent kernel."""
    x = torch.randn(4800, 360, device=DEV)
    w = torch.randn(200, 360, device=DEV)
    ops.gemm_persist_config(0, 65536)
    a = ops.linear_fwd(x, w, 'tf32x3')
    ops.gemm_persist_config(2, 65536)
    b = ops.linear_fwd(x, w, 'tf32x3')
    assert torch.equal(a, b)


@pytest.mark.parametrize('M,N,K', [(3000, 200, 10000), (700, 512, 4096), (128, 64, 2048), (2750, 200, 360)])
def test_split_k_of_long_reductions_in_the_one_tile_kernel(M, N, K):
    """Few output tiles, long reduction (the relation-aware layers' data gradient: [3000, D*H = 10000] x [10000, 200]): the
    K range is split over CTAs and the partial tiles meet in C through vector reductions; the last shape takes the
    unsplit path."""
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    x = torch.randn(M, K, device=DEV, generator=g)
    w = torch.randn(N, K, device=DEV, generator=g) / np.sqrt(K)
    ref = x.double() @ w.double().t()
    y = torch.full((M, N), 7.0, device=DEV)                 # stale contents must not leak into the sums
    for mode, tol in (('tf32x3', 1e-5), ('tf32', 2e-3)):
        ws = ops.weight_prep(w, mode)
        y = ops.linear_fwd(x, w, mode, ws)
        assert _rel(y, ref) <= tol, mode
    dy = torch.randn(M, N, device=DEV, generator=g)          # and as the data gradient: reduction over N
    if N >= 2048 // 4:
        ws = ops.weight_prep(w, 'tf32x3')
        dx = ops.linear_dgrad(dy, w, 'tf32x3', ws)
        assert _rel(dx, dy.double() @ w.double()) <= 1e-5


@pytest.mark.parametrize('cg', (1, 2))
@pytest.mark.parametrize('live', (0, 1, 257, 1000, 3000))
def test_wide_projection_with_a_device_side_row_count(cg, live):
    """gpt_linear_fwd_tf32x3_rows on a wide output (the relation-aware layers' [rows, D*H] projection): the persistent
    kernel deals the tiles round-robin over the clusters and stops at the last live row block -- same bits as the
    one-tile kernel on the live rows, nothing written past the last live block."""
    M, N, K = 3000, 2560, 200
    g = torch.Generator(device=DEV).manual_seed(live + cg)
    x = torch.randn(M, K, device=DEV, generator=g)
    w = torch.randn(N, K, device=DEV, generator=g) / np.sqrt(K)
    ws = ops.weight_prep(w, 'tf32x3')
    bias = torch.randn(N, device=DEV, generator=g)           # added in the epilogue of either kernel
    count = torch.tensor([live], dtype=torch.int32, device=DEV)
    out = []
    for mode in (0, cg):
        ops.gemm_persist_config(mode, 65536)
        y = torch.full((M, N), 7.0, device=DEV)
        ops._call('gpt_linear_fwd_tf32x3_rows', x.data_ptr(), ws.data_ptr(), bias.data_ptr(), y.data_ptr(), M, N, K,
                  count.data_ptr(), ops._stream())
        out.append(y)
    torch.cuda.synchronize()
    assert torch.equal(out[0][:live], out[1][:live])
    if live:
        assert _rel(out[1][:live], x[:live].double() @ w.double().t() + bias.double()) <= 1e-5
    block = 128 * cg
    touched = min(M, (live + block - 1) // block * block)
    assert bool((out[1][touched:] == 7.0).all())


@pytest.mark.parametrize('live', (0, 130, 1000, 2900))
def test_split_k_data_gradient_with_a_device_side_row_count(live):
    """gpt_linear_dgrad_tf32x3_rows (reduction over D*H = 10 000): the number of K ranges is chosen on the device from the
    live row tiles; 3xTF32 tolerance 1e-5."""
    M, N, K = 3000, 10000, 200
    g = torch.Generator(device=DEV).manual_seed(live)
    dy = torch.randn(M, N, device=DEV, generator=g)
    w = torch.randn(N, K, device=DEV, generator=g) / np.sqrt(N)
    ws = ops.weight_prep(w, 'tf32x3')
    count = torch.tensor([live], dtype=torch.int32, device=DEV)
    dx = torch.empty((M, K), device=DEV)
    ops._call('gpt_linear_dgrad_tf32x3_rows', dy.data_ptr(), ws.data_ptr(), dx.data_ptr(), M, N, K, count.data_ptr(),
              ops._stream())
    torch.cuda.synchronize()
    if live:
        assert _rel(dx[:live], dy[:live].double() @ w.double()) <= 1e-5
