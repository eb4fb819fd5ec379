"""CPU: the SOURCE of csrc/gemm_simt.cu (K3 in its fp32 FFMA mode: projection, data gradient, weight gradients incl.
the split-M, accumulating and live-row variants) executed on the host (tests/emu) through ops.linear_* against float64
matmuls.  The `-m gpu` tests of test_gpu_parity.py run the same checks on the device (and the tcgen05 modes, which
cannot be emulated: TMA / tensor memory)."""
import ctypes
import os
import sys

import pytest
import torch

from gcn_over_pruned_trees_b200 import _lib, ops

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
NAMES = ('gpt_linear_fwd_f32', 'gpt_linear_dgrad_f32', 'gpt_linear_wgrad_f32', 'gpt_linear_wgrad_f32_acc',
         'gpt_linear_wgrad_rows_f32')


@pytest.fixture(scope='module', autouse=True)
def emulated():
    import emu_build
    handle = ctypes.CDLL(emu_build.build())
    for name in NAMES:
        getattr(handle, name).argtypes = _lib.SIGNATURES[name]
        getattr(handle, name).restype = ctypes.c_int
    mp = pytest.MonkeyPatch()
    mp.setattr(_lib, '_lib', handle)
    mp.setattr(ops, '_stream', lambda: None)
    yield handle
    mp.undo()


def _rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize('M,N,K', [(1, 1, 1), (50, 200, 360), (700, 64, 85), (130, 37, 19), (1300, 96, 64)])
def test_k3_source_fp32_gemms(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g)
    dy = torch.randn(M, N, generator=g)
    x64, w64, dy64 = x.double(), w.double(), dy.double()
    assert _rel(ops.linear_fwd(x, w), x64 @ w64.t()) < 2e-6
    assert _rel(ops.linear_dgrad(dy, w), dy64 @ w64) < 2e-6
    assert _rel(ops.linear_wgrad(dy, x), dy64.t() @ x64) < 2e-6                # split over M, atomically reduced
    base = torch.randn(N, K, generator=g)
    acc = ops.linear_wgrad(dy, x, out=base.clone(), accumulate=True)           # adds into the caller's buffer
    assert _rel(acc, base.double() + dy64.t() @ x64) < 2e-6


@pytest.mark.parametrize('M,N,K', [(640, 64, 64), (900, 200, 72), (333, 40, 50)])
def test_k3_source_live_row_weight_gradient(M, N, K):
    """gpt_linear_wgrad_rows_f32: only rows whose flag is set are read (their dY is exactly zero otherwise)."""
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, K, generator=g)
    dy = torch.randn(M, N, generator=g)
    flags = (torch.rand(M, generator=g) < 0.3).to(torch.uint8) * (1 + 2 * (torch.rand(M, generator=g) < 0.5).to(torch.uint8))
    dy = dy * flags.ne(0).float().unsqueeze(1)
    out = ops.linear_wgrad(dy, x, out=torch.zeros(N, K), accumulate=True, flags=flags)
    assert _rel(out, dy.double().t() @ x.double()) < 2e-6
