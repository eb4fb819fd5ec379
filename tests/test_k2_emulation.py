"""CPU: the SOURCE of csrc/aggregate.cu (K2: sparse neighbourhood aggregation over the pruned-tree CSR fused with
/denom, bias, ReLU, dropout and -- last layer -- the three max pools; forward and backward) executed on the host
(tests/emu: its shared-memory / cp.async / packed-add PTX accessors replaced by host versions, tests/emu/emu_smem_ops.h;
the cp.async path runs, tensor maps do not exist on the host), on the CSR of the emulated K1 and the projections of the
emulated K3, against the reference's dense formulation (/root/reference/model/gcn.py:260-271, 390-393).  The `-m gpu`
tests of test_gpu_parity.py / test_gpu_fused.py run the same checks on the device, plus the TMA path."""
import ctypes
import os
import sys

import pytest
import torch

from gcn_over_pruned_trees_b200 import _lib, ops, synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
NAMES = ('gpt_prune_csr', 'gpt_linear_fwd_f32', 'gpt_linear_dgrad_f32', 'gpt_linear_wgrad_f32', 'gpt_pool3_fwd',
         'gpt_pool3_bwd', 'gpt_gcn_aggregate_fwd', 'gpt_gcn_aggregate_bwd', 'gpt_gcn_aggregate_bwd_pre',
         'gpt_gcn_aggregate_fwd_pool', 'gpt_gcn_aggregate_fwd_pool_supported', 'gpt_gcn_aggregate_bwd_pool',
         'gpt_gcn_aggregate_bwd_pool_c', 'gpt_gcn_aggregate_bwd_pre_c')


@pytest.fixture(scope='module', autouse=True)
def emulated():
    import emu_build
    handle = ctypes.CDLL(emu_build.build())
    for name in NAMES:
        getattr(handle, name).argtypes = _lib.SIGNATURES[name]
        getattr(handle, name).restype = ctypes.c_int
    mp = pytest.MonkeyPatch()
    mp.setattr(_lib, '_lib', handle)
    mp.setattr(ops, '_dev', lambda t, dtype, name: t.contiguous() if t.dtype == dtype else (_ for _ in ()).throw(
        TypeError('%s must be %s' % (name, dtype))))
    mp.setattr(ops, '_stream', lambda: None)
    yield handle
    mp.undo()


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _csr_of(batch, k):
    return ops.prune_csr(batch[5], batch[6], batch[7], batch[4], batch[1], k)


def _dense_layer(x, w, b, adj, mask=None):
    a = (adj != 0).double()
    denom = a.sum(2, keepdim=True) + 1
    z = ((a.bmm(x) @ w.t() + b) + (x @ w.t() + b)) / denom
    out = torch.relu(z)
    return out if mask is None else out * mask


@pytest.mark.parametrize('k', (-1, 1))
@pytest.mark.parametrize('H,K,vec', [(200, 64, 0), (64, 40, 2), (200, 48, 4), (30, 17, 0)])
def test_k2_source_layer_forward_backward_vs_dense(k, H, K, vec):
    batch = synth.make_batch(20 + k, batch_size=6)
    csr = _csr_of(batch, k)
    B, T = batch[0].shape
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, T, K, generator=g, requires_grad=True)
    w = (torch.randn(H, K, generator=g) / K ** 0.5).requires_grad_()
    b = torch.randn(H, generator=g).requires_grad_()
    mask = (torch.rand(B, T, H, generator=g) < 0.5).float() * 2.0
    gout = torch.randn(B, T, H, generator=g)
    observable = (csr.flags != 0).unsqueeze(2).double()          # rows that can reach the logits
    out = ops.gcn_layer(x, w, b, csr, drop_mask=mask)
    if vec:
        y = ops.linear_fwd(x.detach().view(B * T, K), w.detach())
        assert torch.equal(ops.aggregate_fwd(y, csr, b.detach(), drop_mask=mask, force_vec=vec), out.detach())
    (out * gout * observable.float()).sum().backward()
    adj = csr.to_dense()
    xd, wd, bd = (t.detach().double().requires_grad_() for t in (x, w, b))
    ref = _dense_layer(xd, wd, bd, adj, mask.double())
    (ref * gout.double() * observable).sum().backward()
    assert _rel(out.detach().double() * observable, ref.detach() * observable) < 1e-5
    assert torch.all(out.detach()[csr.flags == 0] == 0)
    assert _rel(x.grad, xd.grad) < 1e-5
    assert _rel(w.grad, wd.grad) < 1e-5
    assert _rel(b.grad, bd.grad) < 1e-5


def test_k2_source_no_adj_ablation():
    batch = synth.make_batch(31, batch_size=4)
    csr = _csr_of(batch, 1)
    B, T = batch[0].shape
    g = torch.Generator().manual_seed(1)
    x, w, b = torch.randn(B, T, 32, generator=g), torch.randn(48, 32, generator=g), torch.randn(48, generator=g)
    out = ops.gcn_layer(x, w, b, csr, use_adj=False)
    ref = torch.relu((x @ w.t() + 2 * b) / csr.denom.unsqueeze(2)) * (csr.flags != 0).unsqueeze(2)
    assert _rel(out, ref) < 1e-5


def test_k2_source_philox_dropout_statistics_and_backward_consistency():
    batch = synth.make_batch(33, batch_size=12)
    csr = _csr_of(batch, -1)
    B, T = batch[0].shape
    H = 72
    g = torch.Generator().manual_seed(2)
    y = torch.rand(B * T, H, generator=g) + 0.5                    # strictly positive pre-activations
    bias = torch.zeros(H)
    rng = torch.tensor([1234, 1], dtype=torch.int64)
    base = ops.aggregate_fwd(y, csr, bias)
    for p in (0.5, 0.1):
        o1 = ops.aggregate_fwd(y, csr, bias, drop_p=p, rng_state=rng, subseq=0)
        assert torch.equal(o1, ops.aggregate_fwd(y, csr, bias, drop_p=p, rng_state=rng, subseq=0, force_vec=2))
        live = base > 0
        assert abs((o1 > 0)[live].float().mean().item() - (1 - p)) < 0.02
        scale = 65536.0 / (65536 - round(p * 65536))              # keep-probability is quantised to 16 bits
        assert _rel(o1[o1 > 0], base[o1 > 0] * scale) < 1e-6
        o2 = ops.aggregate_fwd(y, csr, bias, drop_p=p, rng_state=rng, subseq=1)
        assert not torch.equal(o1 > 0, o2 > 0)
        gout = torch.randn(B, T, H, generator=g)
        dy1, db1 = ops.aggregate_bwd(gout, o1, csr, drop_p=p)
        dy2, db2 = ops.aggregate_bwd(gout, o1, csr, drop_mask=(o1 > 0).float() * scale)
        assert _rel(dy1, dy2) < 1e-6 and _rel(db1, db2) < 1e-5
        o4, act = ops.aggregate_fwd(y, csr, bias, drop_p=p, rng_state=rng, subseq=0, want_act=True)
        assert torch.equal(o4, o1)
        dy3, db3 = ops.aggregate_bwd(gout, None, csr, drop_p=p, act=act)  # driven by the 1-bit activation mask
        assert torch.equal(dy3, dy1) and _rel(db3, db1) < 1e-5


def test_k2_source_last_layer_fused_with_the_three_max_pools_both_directions():
    """gpt_gcn_aggregate_fwd_pool / _bwd_pool == K2 then K4 (forward) and K4-backward then K2-backward, bit for bit."""
    batch = synth.make_batch(35, batch_size=6)
    csr = _csr_of(batch, 1)
    B, T = batch[0].shape
    H = 64
    assert ops.aggregate_pool_ok(B, T, H)
    g = torch.Generator().manual_seed(4)
    y = torch.randn(B * T, H, generator=g)
    bias = torch.randn(H, generator=g)
    pooled, argmax, act, out = ops.aggregate_fwd_pool(y, csr, bias, want_out=True)
    out_ref, act_ref = ops.aggregate_fwd(y, csr, bias, want_act=True)
    pooled_ref, argmax_ref = ops.pool3_fwd(out_ref, csr, 0)
    assert torch.equal(out, out_ref) and torch.equal(pooled, pooled_ref) and torch.equal(argmax, argmax_ref)
    assert torch.equal(act, act_ref)
    dpooled = torch.randn(B, 3 * H, generator=g)
    dy = ops.aggregate_bwd_pool(dpooled, argmax, act, csr, H)
    dh = ops.pool3_bwd(dpooled, argmax_ref, csr, 0, H)
    dy_ref, _ = ops.aggregate_bwd(dh, None, csr, act=act_ref)
    assert torch.equal(dy, dy_ref)


def test_k2_source_fused_pool_forward_in_128_thread_ctas(monkeypatch):
    """The 128-thread form of the fused-pool forward (chosen when the 256-thread form would need a second wave of CTAs,
    e.g. 50 sentences x 7 slices) == the 256-thread form, bit for bit, ragged widths included."""
    batch = synth.make_batch(36, batch_size=9)
    csr = _csr_of(batch, 1)
    B, T = batch[0].shape
    g = torch.Generator().manual_seed(5)
    for H in (64, 200, 72):
        y = torch.randn(B * T, H, generator=g)
        bias = torch.randn(H, generator=g)
        monkeypatch.setenv('GPT_AGG_POOL_NT', '256')
        ref = ops.aggregate_fwd_pool(y, csr, bias, want_out=True)
        monkeypatch.setenv('GPT_AGG_POOL_NT', '128')
        got = ops.aggregate_fwd_pool(y, csr, bias, want_out=True)
        for a, b in zip(got, ref):
            assert torch.equal(a, b)


@pytest.mark.parametrize('k', (-1, 1))
def test_k2_source_persistent_double_buffered_path_512_tokens(k, monkeypatch):
    """Large sentence tiles (the shape the roofline number is quoted on): the 512-thread CTA walks the sentence's column
    slices with two buffers, the next slice in flight while the current one is gathered.  GPT_AGG_SPLIT=1 gives one
    CTA per sentence, as a full machine would (B >= 296)."""
    monkeypatch.setenv('GPT_AGG_SPLIT', '1')
    B, T, H = 2, 512, 96
    batch = synth.make_batch(900 + k, batch_size=B, fixed_len=T)
    csr = _csr_of(batch, k)
    assert int((csr.err & ops.TREE_ERR_FATAL).sum()) == 0
    g = torch.Generator().manual_seed(3)
    y = torch.randn(B * T, H, generator=g)
    bias = torch.randn(H, generator=g)
    gout = torch.randn(B, T, H, generator=g)
    rng = torch.tensor([77, 2], dtype=torch.int64)
    out, act = ops.aggregate_fwd(y, csr, bias, want_act=True)
    assert torch.equal(out, ops.aggregate_fwd(y, csr, bias, force_vec=1))          # one slice per CTA, no pipelining
    adj = (csr.to_dense() != 0).double()
    ys = y.view(B, T, H).double()
    ref = torch.relu((adj.bmm(ys) + ys + 2 * bias.double()) / csr.denom.double().unsqueeze(2)) * \
        (csr.flags != 0).unsqueeze(2)
    assert _rel(out, ref) < 1e-5
    dy, db = ops.aggregate_bwd(gout, None, csr, act=act)
    dy_small, db_small = ops.aggregate_bwd(gout, out, csr, force_vec=1)
    assert torch.equal(dy, dy_small) and _rel(db, db_small) < 1e-5
    # what the training step launches at this shape: in-kernel dropout 0.5 (1 random bit per element) + activation mask
    o1, act1 = ops.aggregate_fwd(y, csr, bias, drop_p=0.5, rng_state=rng, subseq=0, want_act=True)
    assert torch.equal(o1, ops.aggregate_fwd(y, csr, bias, drop_p=0.5, rng_state=rng, subseq=0, force_vec=1))
    live = out > 0
    assert abs((o1 > 0)[live].float().mean().item() - 0.5) < 0.02
    assert torch.equal(o1[o1 > 0], out[o1 > 0] * 2.0)
