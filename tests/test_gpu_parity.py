"""GPU parity tests (run on the B200 box with ``-m gpu``): every CUDA entry point, through the C ABI, against the
oracle and against the committed reference outputs (tests/golden/*.npz).

Tolerances (stated here once):
  * adjacency: bit-exact (integer work)
  * fp32 path: logits / loss <= 1e-5 relative (max|d| / max|ref|), gradients <= 1e-4 relative per tensor
"""
import numpy as np
import pytest
import torch

import cases
import weights
from gcn_over_pruned_trees_b200 import ops, synth, _lib
from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer
from oracle import gcn_oracle, tree_oracle

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(params=['warp_per_sentence', 'cta_per_sentence'])
def k1_form(request, monkeypatch):
    """Both forms of K1 (one warp / one CTA per sentence) on every K1 case; the library reads the threshold per call."""
    monkeypatch.setenv('GPT_K1_BLOCK_MIN_T', '1' if request.param == 'cta_per_sentence' else '1000000')
    return request.param


def _csr_of(batch, k, dataset='tacred'):
    off = 4 if dataset == 'tacred' else 3
    deprel, head, subj_pos, obj_pos = [t.to(DEV) for t in batch[off:off + 4]]
    return ops.prune_csr(head, subj_pos, obj_pos, deprel, batch[1].to(DEV), k)


def _oracle_adj(batch, k):
    lens = synth.batch_lengths(batch).numpy()
    return tree_oracle.batch_adjacency(batch[5].numpy(), batch[6].numpy(), batch[7].numpy(), batch[4].numpy(), lens,
                                       k, batch[0].shape[1])


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


# ---------------------------------------------------------------- K1 ------------------------------------------

@pytest.mark.parametrize('split', cases.SPLITS)
def test_k1_bundled_sample_matches_reference(golden_adj, split, k1_form):
    batch = cases.batch_from_npz(golden_adj, split)
    for k in cases.PRUNE_KS:
        csr = _csr_of(batch, k)
        assert int(csr.err.abs().sum()) == 0
        got = csr.to_dense().numpy()
        want = golden_adj['%s/adj_k%d' % (split, k)].astype(np.float32)
        assert np.array_equal(got, want), (split, k)
        # denom = rowsum(adj != 0) + 1 ; in-tree flag = (rowsum + colsum) != 0  (gcn.py:261-262)
        nz = want != 0
        assert np.array_equal(csr.denom.cpu().numpy(), nz.sum(2) + 1.0)
        assert np.array_equal((csr.flags.cpu().numpy() & 1) != 0, (nz.sum(2) + nz.sum(1)) != 0)
        assert np.array_equal(csr.lens.cpu().numpy(), synth.batch_lengths(batch).numpy())


@pytest.mark.parametrize('seed', cases.SYNTH_ADJ_SEEDS + (1, 2, 3))
def test_k1_synthetic_matches_oracle(seed, k1_form):
    batch = synth.make_batch(seed, batch_size=50)
    for k in cases.PRUNE_KS + (7,):
        csr = _csr_of(batch, k)
        got = csr.to_dense().numpy()
        assert np.array_equal(got, _oracle_adj(batch, k)), (seed, k)
        # columns ascending inside every row (canonical CSR)
        rp, col = csr.rowptr.cpu().numpy(), csr.col.cpu().numpy()
        for b in range(0, 50, 7):
            for t in range(batch[0].shape[1]):
                seg = col[b, rp[b, t]:rp[b, t + 1]]
                assert np.all(np.diff(seg) > 0)


def test_k1_subj_obj_flags(k1_form):
    batch = synth.make_batch(4, batch_size=20)
    csr = _csr_of(batch, 1)
    f = csr.flags.cpu().numpy()
    assert np.array_equal((f & 2) != 0, batch[6].numpy() == 0)
    assert np.array_equal((f & 4) != 0, batch[7].numpy() == 0)


@pytest.mark.parametrize('name', sorted(cases.EDGE_TREES))
def test_k1_edge_trees(golden_adj, name, k1_form):
    head, subj, obj, deprel = cases.EDGE_TREES[name]
    n, width = len(head), len(head) + 3           # padded on purpose
    pad = lambda a, fill=0: torch.tensor([list(a) + [fill] * (width - n)], dtype=torch.int64, device=DEV)
    masks = torch.tensor([[False] * n + [True] * (width - n)], device=DEV)
    sp = torch.from_numpy(cases.positions(subj, n, width=width))[None].to(DEV)
    op = torch.from_numpy(cases.positions(obj, n, width=width))[None].to(DEV)
    for k in cases.PRUNE_KS:
        csr = ops.prune_csr(pad(head), sp, op, pad(deprel), masks, k)
        assert (int(csr.err[0]) & ops.TREE_ERR_FATAL) == 0, (name, k, int(csr.err[0]))
        got = csr.to_dense().numpy()[0, :n, :n]
        assert np.array_equal(got, golden_adj['edge/%s/k%d' % (name, k)].astype(np.float32)), (name, k)
        assert csr.to_dense().numpy()[0, n:, :].sum() == 0


def test_k1_512_token_sentences(golden_adj, k1_form):
    import hashlib
    batch = synth.make_batch(900, batch_size=6, fixed_len=512)
    for k in (-1, 1):
        got = _csr_of(batch, k).to_dense().numpy()
        sha = hashlib.sha256(got.astype(np.uint8).tobytes()).hexdigest()[:16]
        assert sha == bytes(golden_adj['synth512/k%d/sha' % k]).decode()


def test_k1_malformed_trees_are_flagged_not_hung(k1_form):
    def run(head, subj, obj, k, deprel=None):
        n = len(head)
        t = lambda a: torch.tensor([a], dtype=torch.int64, device=DEV)
        deprel = deprel or [5] * n
        csr = ops.prune_csr(t(head), torch.from_numpy(cases.positions(subj, n))[None].to(DEV),
                            torch.from_numpy(cases.positions(obj, n))[None].to(DEV), t(deprel),
                            torch.zeros((1, n), dtype=torch.bool, device=DEV), k)
        return int(csr.err[0]), csr
    e, csr = run([2, 3, 1], [0], [2], 1)
    assert e & 4 and csr.to_dense().sum() == 0                       # cycle (reference: infinite loop)
    assert run([2, 3, 1], [0], [2], -1)[0] & (2 | 4)                 # no root
    assert run([0, 0], [0], [1], 0)[0] & 16                          # entities under different roots
    assert run([0, 1], [], [1], 0)[0] & 8                            # empty subject span
    assert run([0, 5], [0], [1], 0)[0] & 1                           # head out of range
    assert run([0, 1], [0], [1], 0, deprel=[11, 250])[0] & 32        # deprel does not fit uint8 + 42
    assert run([0, 1, 1, 2], [3], [2], 1, deprel=[11, 0, 5, 0])[0] == 64   # warning only
    with pytest.raises(_lib.GptError):
        csr.err.fill_(4)
        csr.check()


def test_no_cpu_fallback():
    batch = synth.make_batch(0, batch_size=4)
    with pytest.raises(_lib.GptError):
        ops.prune_csr(batch[5], batch[6], batch[7], batch[4], batch[1], 1)


# ---------------------------------------------------------------- K3 ------------------------------------------

@pytest.mark.parametrize('M,N,K', [(1, 1, 1), (50, 200, 360), (2750, 200, 360), (1000, 200, 200), (777, 64, 85),
                                    (4096, 512, 360), (130, 37, 19)])
def test_k3_fp32_gemms(M, N, K):
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    x = torch.randn(M, K, device=DEV, generator=g)
    w = torch.randn(N, K, device=DEV, generator=g)
    dy = torch.randn(M, N, device=DEV, generator=g)
    x64, w64, dy64 = x.double(), w.double(), dy.double()
    assert _rel(ops.linear_fwd(x, w).cpu(), (x64 @ w64.t()).cpu()) < 2e-6
    assert _rel(ops.linear_dgrad(dy, w).cpu(), (dy64 @ w64).cpu()) < 2e-6
    assert _rel(ops.linear_wgrad(dy, x).cpu(), (dy64.t() @ x64).cpu()) < 2e-6


# ---------------------------------------------------------------- K2 ------------------------------------------

def _dense_layer(x, w, b, adj, mask=None):
    a = (adj != 0).double()
    denom = a.sum(2, keepdim=True) + 1
    z = ((a.bmm(x) @ w.t() + b) + (x @ w.t() + b)) / denom
    out = torch.relu(z)
    return out if mask is None else out * mask


@pytest.mark.parametrize('k', (-1, 0, 1, 2))
@pytest.mark.parametrize('H,K,vec', [(200, 360, 0), (200, 200, 1), (64, 40, 2), (200, 64, 4), (30, 17, 0)])
def test_k2_layer_forward_backward_vs_dense(k, H, K, vec):
    batch = synth.make_batch(20 + k, batch_size=24)
    csr = _csr_of(batch, k)
    B, T = batch[0].shape
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn(B, T, K, device=DEV, generator=g, requires_grad=True)
    w = (torch.randn(H, K, device=DEV, generator=g) / K ** 0.5).requires_grad_()
    b = torch.randn(H, device=DEV, generator=g).requires_grad_()
    mask = (torch.rand(B, T, H, device=DEV, generator=g) < 0.5).float() * 2.0
    gout = torch.randn(B, T, H, device=DEV, generator=g)
    observable = (csr.flags != 0).unsqueeze(2).double()          # rows that can reach the logits

    out = ops.gcn_layer(x, w, b, csr, drop_mask=mask)
    if vec:
        y = ops.linear_fwd(x.detach().view(B * T, K), w.detach())
        out_v = ops.aggregate_fwd(y, csr, b.detach(), drop_mask=mask, force_vec=vec)
        assert torch.equal(out_v, out.detach())
    (out * gout * observable.float()).sum().backward()

    adj = csr.to_dense().to(DEV)
    xd, wd, bd = (t.detach().double().requires_grad_() for t in (x, w, b))
    ref = _dense_layer(xd, wd, bd, adj, mask.double())
    (ref * gout.double() * observable).sum().backward()
    assert _rel((out.detach().double() * observable).cpu(), (ref.detach() * observable).cpu()) < 1e-5
    assert torch.all(out.detach()[csr.flags == 0] == 0)
    assert _rel(x.grad.cpu(), xd.grad.cpu()) < 1e-5
    assert _rel(w.grad.cpu(), wd.grad.cpu()) < 1e-5
    assert _rel(b.grad.cpu(), bd.grad.cpu()) < 1e-5


@pytest.mark.parametrize('k', (-1, 1))
def test_k2_persistent_double_buffered_path_512_tokens(k):
    # large sentence tiles: the CTA walks the sentence's column slices with two buffers (B >= 148 forces 2 slices/CTA)
    B, T, H = 160, 512, 128
    batch = synth.make_batch_torch(11, B, T, device=DEV)
    csr = ops.prune_csr(batch[5], batch[6], batch[7], batch[4], batch[1], k)
    assert int((csr.err & ops.TREE_ERR_FATAL).sum()) == 0
    g = torch.Generator(device=DEV).manual_seed(3)
    y = torch.randn(B * T, H, device=DEV, generator=g)
    bias = torch.randn(H, device=DEV, generator=g)
    gout = torch.randn(B, T, H, device=DEV, generator=g)
    out, act = ops.aggregate_fwd(y, csr, bias, want_act=True)
    ref_small = ops.aggregate_fwd(y, csr, bias, force_vec=1)           # one slice per CTA, no pipelining
    assert torch.equal(out, ref_small)
    # independent check of a few sentences against the dense formulation
    sel = [0, 77, 159]
    adj = torch.zeros(len(sel), T, T, device=DEV)
    rp, col = csr.rowptr.cpu(), csr.col.cpu()
    for n, b in enumerate(sel):
        counts = (rp[b, 1:] - rp[b, :-1]).long()
        rows = torch.repeat_interleave(torch.arange(T), counts)
        adj[n, rows, col[b, :int(rp[b, T])].long()] = 1.0
    ys = y.view(B, T, H)[sel].double()
    dn = csr.denom[sel].double().unsqueeze(2)
    ref = torch.relu((adj.double().bmm(ys) + ys + 2 * bias.double()) / dn) * (csr.flags[sel] != 0).unsqueeze(2)
    assert _rel(out[sel].cpu(), ref.cpu()) < 1e-5
    dy, db = ops.aggregate_bwd(gout, None, csr, act=act)
    dy_small, db_small = ops.aggregate_bwd(gout, out, csr, force_vec=1)
    assert torch.equal(dy, dy_small) and _rel(db.cpu(), db_small.cpu()) < 1e-5


def test_k2_no_adj_ablation():
    batch = synth.make_batch(31, batch_size=8)
    csr = _csr_of(batch, 1)
    B, T = batch[0].shape
    x = torch.randn(B, T, 32, device=DEV)
    w = torch.randn(48, 32, device=DEV)
    b = torch.randn(48, device=DEV)
    out = ops.gcn_layer(x, w, b, csr, use_adj=False)
    ref = torch.relu((x @ w.t() + 2 * b) / csr.denom.unsqueeze(2)) * (csr.flags != 0).unsqueeze(2)
    assert _rel(out.cpu(), ref.cpu()) < 1e-5


def test_k2_philox_dropout_statistics_and_backward_consistency():
    batch = synth.make_batch(33, batch_size=50)
    csr = _csr_of(batch, -1)
    B, T = batch[0].shape
    H = 200
    y = torch.rand(B * T, H, device=DEV) + 0.5                     # strictly positive pre-activations
    bias = torch.zeros(H, device=DEV)
    rng = torch.tensor([1234, 1], dtype=torch.int64, device=DEV)
    base = ops.aggregate_fwd(y, csr, bias)
    for p in (0.5, 0.1):
        o1 = ops.aggregate_fwd(y, csr, bias, drop_p=p, rng_state=rng, subseq=0)
        o1b = ops.aggregate_fwd(y, csr, bias, drop_p=p, rng_state=rng, subseq=0, force_vec=4)
        assert torch.equal(o1, o1b)                                # pattern independent of the slicing
        live = base > 0
        keep = (o1 > 0)[live].float().mean().item()
        assert abs(keep - (1 - p)) < 0.01
        scale = 65536.0 / (65536 - round(p * 65536))              # keep-probability is quantised to 16 bits
        assert _rel(o1[o1 > 0].cpu(), (base[o1 > 0] * scale).cpu()) < 1e-6
        o2 = ops.aggregate_fwd(y, csr, bias, drop_p=p, rng_state=rng, subseq=1)
        rng2 = torch.tensor([1234, 2], dtype=torch.int64, device=DEV)
        o3 = ops.aggregate_fwd(y, csr, bias, drop_p=p, rng_state=rng2, subseq=0)
        assert not torch.equal(o1 > 0, o2 > 0) and not torch.equal(o1 > 0, o3 > 0)
        # backward with the in-kernel mask == backward with the same mask given explicitly
        gout = torch.randn(B, T, H, device=DEV)
        dy1, db1 = ops.aggregate_bwd(gout, o1, csr, drop_p=p)
        explicit = (o1 > 0).float() * scale
        dy2, db2 = ops.aggregate_bwd(gout, o1, csr, drop_mask=explicit)
        assert _rel(dy1.cpu(), dy2.cpu()) < 1e-6 and _rel(db1.cpu(), db2.cpu()) < 1e-5
        # ... and == backward driven by the forward's 1-bit activation mask instead of `out`
        o4, act = ops.aggregate_fwd(y, csr, bias, drop_p=p, rng_state=rng, subseq=0, want_act=True)
        assert torch.equal(o4, o1)
        dy3, db3 = ops.aggregate_bwd(gout, None, csr, drop_p=p, act=act)
        assert torch.equal(dy3, dy1) and _rel(db3.cpu(), db1.cpu()) < 1e-5


# ---------------------------------------------------------------- K4 ------------------------------------------

@pytest.mark.parametrize('kind', ('max', 'avg', 'sum'))
@pytest.mark.parametrize('H', (200, 37))
def test_k4_pool3_vs_reference_pool(kind, H):
    batch = synth.make_batch(41, batch_size=16)
    csr = _csr_of(batch, 1)
    B, T = batch[0].shape
    h = (torch.rand(B, T, H, device=DEV) + 0.01).requires_grad_()      # no ties: argmax is unique
    gout = torch.randn(B, 3 * H, device=DEV)
    got = ops.pool3(h, csr, kind)
    (got * gout).sum().backward()
    hd = h.detach().clone().requires_grad_()
    m = [csr.pool_mask(), batch[6].to(DEV).ne(0).unsqueeze(2), batch[7].to(DEV).ne(0).unsqueeze(2)]
    ref = torch.cat([gcn_oracle.masked_pool(hd, mi, kind) for mi in m], dim=1)
    (ref * gout).sum().backward()
    assert _rel(got.detach().cpu(), ref.detach().cpu()) < 1e-6
    assert _rel(h.grad.cpu(), hd.grad.cpu()) < 1e-6


def test_k4_fully_masked_pool_is_minus_1e12():
    # same-token subject/object -> singleton tree -> empty adjacency -> h_out = -1e12
    head = torch.tensor([[2, 0, 2, 3]], device=DEV)
    pos = torch.from_numpy(cases.positions([3], 4))[None].to(DEV)
    csr = ops.prune_csr(head, pos, pos, torch.tensor([[5, 11, 6, 7]], device=DEV),
                        torch.zeros((1, 4), dtype=torch.bool, device=DEV), 1)
    h = torch.rand(1, 4, 8, device=DEV)
    out = ops.pool3(h, csr, 'max')
    assert torch.all(out[:, :8] == -1e12)
    assert torch.equal(out[:, 8:16], h[:, 3]) and torch.equal(out[:, 16:], h[:, 3])


# ---------------------------------------------------------------- whole model ---------------------------------

def _setup(golden_adj, name, gemm_mode='fp32'):
    over, source, wseed = cases.MODEL_CASES[name]
    over = dict(over, gemm_mode=gemm_mode)
    if source[0] == 'split':
        batch = cases.batch_from_npz(golden_adj, source[1])
        over = dict(over, vocab_size=int(golden_adj['vocab_size']))
    else:
        batch = synth.make_batch(source[1], batch_size=source[2], vocab_size=over['vocab_size'],
                                 num_class=over.get('num_class', 42), dataset=over.get('dataset', 'tacred'))
    opt = synth.tacred_opt(**over)
    state = {k: torch.from_numpy(v) for k, v in weights.make_state(opt, wseed).items()}
    trainer = GCNTrainer(dict(opt, cuda=True))
    trainer.model.load_state_dict(state)
    oracle = gcn_oracle.DenseClassifier(opt)
    oracle.load_state_dict(state)
    return opt, batch, trainer, oracle


@pytest.mark.parametrize('gemm_mode', ('fp32', 'tf32x3'))
@pytest.mark.parametrize('name', sorted(cases.MODEL_CASES))
def test_model_eval_matches_reference_outputs(golden_adj, golden_model, name, gemm_mode):
    opt, batch, trainer, _ = _setup(golden_adj, name, gemm_mode)
    trainer.model.eval()
    with torch.no_grad():
        inputs = [t.to(DEV) for t in batch[:-2]]
        logits, h_out = trainer.model(inputs)
        loss = trainer.update(batch)
    assert _rel(logits.cpu(), golden_model['%s/logits' % name]) <= 1e-5
    assert _rel(h_out.cpu(), golden_model['%s/h_out' % name]) <= 1e-5
    assert abs(loss.item() - float(golden_model['%s/eval_loss' % name])) <= 1e-5 * abs(loss.item())
    preds, probs, ploss = trainer.predict(batch)
    assert preds == golden_model['%s/pred' % name].tolist()
    assert _rel(np.asarray(probs), golden_model['%s/probs' % name]) <= 1e-5
    assert abs(ploss - float(golden_model['%s/predict_loss' % name])) <= 1e-5 * abs(ploss)


@pytest.mark.parametrize('gemm_mode', ('fp32', 'tf32x3'))
@pytest.mark.parametrize('name', cases.GRAD_CASES)
def test_model_train_grads_match_oracle_with_injected_masks(golden_adj, name, gemm_mode):
    opt, batch, trainer, oracle = _setup(golden_adj, name, gemm_mode)
    B, T = batch[0].shape
    g = torch.Generator().manual_seed(99)
    in_dim = opt['emb_dim'] + opt['pos_dim'] + (opt['ner_dim'] if opt['dataset'] == 'tacred' else 0)

    def drop(shape, p):
        return (torch.rand(shape, generator=g) >= p).float() / (1 - p)
    masks = {'in': drop((B, T, in_dim), opt['input_dropout'])}
    if opt['rnn']:
        masks['rnn'] = drop((B, T, 2 * opt['rnn_hidden']), opt['rnn_dropout'])
    for l in range(opt['num_layers'] - 1):
        masks['gcn%d' % l] = drop((B, T, opt['hidden_dim']), opt['gcn_dropout'])

    oracle.train()
    ref_loss, _ = oracle.loss(batch, masks)
    ref_loss.backward()

    trainer.model.train()
    trainer.model.gcn_model.gcn.injected_masks = {k: v.to(DEV) for k, v in masks.items()}
    loss = trainer.update(batch)
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    ref_grads = {k: p.grad for k, p in oracle.named_parameters()}
    checked = 0
    for key, p in trainer.model.named_parameters():
        rg = ref_grads[key]
        if rg is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, key
            continue
        assert _rel(p.grad.cpu(), rg) <= 1e-4, key
        checked += 1
    assert checked >= 8


@pytest.mark.parametrize('fast', (True, False))
def test_train_loop_step_runs_and_lowers_loss(fast):
    """train.py:213-227 with dropout on; fast=True: update / backward / step replay captured graphs (engine.FastUpdate)."""
    opt = synth.tacred_opt(vocab_size=500, cuda=True)
    torch.manual_seed(0)
    trainer = GCNTrainer(opt)
    trainer.fast_update = fast
    batch = synth.make_batch(7, batch_size=50, vocab_size=500)
    trainer.model.train()
    losses = []
    for _ in range(12):                                   # train.py:213-227
        trainer.optimizer.zero_grad()
        loss = trainer.update(batch)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(trainer.model.parameters(), opt['max_grad_norm'])
        trainer.optimizer.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]


def test_checkpoint_round_trip(tmp_path):
    opt = synth.tacred_opt(vocab_size=300, cuda=True)
    a = GCNTrainer(opt)
    f = str(tmp_path / 'ckpt.pt')
    a.save(f, 1)
    from gcn_over_pruned_trees_b200 import torch_utils
    cfg = torch_utils.load_config(f)
    b = GCNTrainer(cfg)
    b.load(f)
    batch = synth.make_batch(3, batch_size=10, vocab_size=300)
    pa, _, la = a.predict(batch)
    pb, _, lb = b.predict(batch)
    assert pa == pb and la == lb
    sd = b.model.state_dict()
    assert sd['gcn_model.emb.weight'].data_ptr() == sd['gcn_model.gcn.emb.weight'].data_ptr()


# ---------------------------------------------------------------- engine --------------------------------------

def test_graphed_train_step_matches_eager_steps():
    # dropout off so that the two runs are comparable; same init, same batches
    over = dict(vocab_size=700, cuda=True, input_dropout=0.0, gcn_dropout=0.0)
    batches = [synth.make_batch(50 + i, batch_size=50, vocab_size=700, pad_to=64) for i in range(3)]

    def run(graphed):
        torch.manual_seed(5)
        tr = GCNTrainer(synth.tacred_opt(**over))
        tr.fast_update = False              # the eager side of this comparison is the per-op autograd path
        tr.model.train()
        losses = []
        for step in range(9):
            b = batches[step % 3]
            if graphed:
                losses.append(float(tr.train_step(b)))
            else:
                tr.optimizer.zero_grad()
                loss = tr.update(b)
                loss.backward()
                torch.nn.utils.clip_grad_norm_(tr.model.parameters(), tr.opt['max_grad_norm'])
                tr.optimizer.step()
                losses.append(loss.item())
        return losses, {k: v.detach().cpu() for k, v in tr.model.state_dict().items()}, tr

    l0, s0, _ = run(False)
    l1, s1, tr = run(True)
    assert tr._graphed.replays >= 5                       # steps 4.. of the single shape are graph replays
    assert np.allclose(l0, l1, rtol=2e-5)
    for k in s0:
        assert _rel(s1[k], s0[k]) < 5e-5, k


def test_graphed_train_step_with_dropout_trains():
    torch.manual_seed(0)
    tr = GCNTrainer(synth.tacred_opt(vocab_size=500, cuda=True))
    tr.model.train()
    batch = synth.make_batch(7, batch_size=50, vocab_size=500)
    losses = [float(tr.train_step(batch)) for _ in range(30)]
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < np.mean(losses[:5])
    assert len(set(losses[10:20])) > 1                    # dropout stream advances between replays


# ---------------------------------------------------------------- K5 ------------------------------------------

@pytest.mark.parametrize('dataset', ('tacred', 'semeval'))
def test_k5_embed_concat_vs_torch(dataset):
    V, E, Dp, Dn = 300, 300, 30, 30 if dataset == 'tacred' else 0
    batch = synth.make_batch(61, batch_size=20, vocab_size=V, dataset=dataset)
    words, pos = batch[0].to(DEV), batch[2].to(DEV)
    ner = batch[3].to(DEV) if dataset == 'tacred' else None
    g = torch.Generator(device=DEV).manual_seed(1)
    tabs = [torch.randn(V, E, device=DEV, generator=g).requires_grad_(),
            torch.randn(47, Dp, device=DEV, generator=g).requires_grad_(),
            torch.randn(15, 30, device=DEV, generator=g).requires_grad_() if Dn else None]
    x = ops.embed_concat(words, pos, ner, tabs[0], tabs[1], tabs[2])
    r = torch.randn(x.shape, device=DEV, generator=g)
    (x * r).sum().backward()
    ref_t = [t.detach().clone().requires_grad_() if t is not None else None for t in tabs]
    parts = [torch.nn.functional.embedding(words, ref_t[0], padding_idx=0), torch.nn.functional.embedding(pos, ref_t[1])]
    if Dn:
        parts.append(torch.nn.functional.embedding(ner, ref_t[2]))
    xr = torch.cat(parts, 2)
    (xr * r).sum().backward()
    assert torch.equal(x, xr)
    for a, b in zip(tabs, ref_t):
        if a is not None:
            assert _rel(a.grad.cpu(), b.grad.cpu()) < 1e-5


def test_k5_backward_over_many_rows_collects_the_small_tables_in_shared_memory():
    """>= 64 K token rows: the multi-row backward kernel (POS / NER gradients summed in shared memory, flushed once per
    CTA) against torch, including POS ids beyond the 64 rows the shared table holds."""
    V, E, Dp, Dn, n = 500, 300, 30, 30, 70_001
    g = torch.Generator(device=DEV).manual_seed(5)
    words = torch.randint(0, V, (1, n), device=DEV, generator=g)
    pos = torch.randint(0, 80, (1, n), device=DEV, generator=g)           # ids 64..79 take the direct path
    ner = torch.randint(0, 15, (1, n), device=DEV, generator=g)
    tabs = [torch.randn(V, E, device=DEV, generator=g).requires_grad_(),
            torch.randn(80, Dp, device=DEV, generator=g).requires_grad_(),
            torch.randn(15, Dn, device=DEV, generator=g).requires_grad_()]
    x = ops.embed_concat(words, pos, ner, tabs[0], tabs[1], tabs[2])
    r = torch.randn(x.shape, device=DEV, generator=g)
    (x * r).sum().backward()
    ref_t = [t.detach().clone().requires_grad_() for t in tabs]
    xr = torch.cat([torch.nn.functional.embedding(words, ref_t[0], padding_idx=0),
                    torch.nn.functional.embedding(pos, ref_t[1]), torch.nn.functional.embedding(ner, ref_t[2])], 2)
    (xr * r).sum().backward()
    assert torch.equal(x, xr)
    for a, b in zip(tabs, ref_t):
        assert _rel(a.grad.double().cpu(), b.grad.double().cpu()) < 2e-5      # thousands of fp32 additions per entry


def test_k5_dropout_mask_is_replayed_in_backward_and_topn_freezes_rows():
    V, E = 200, 64
    batch = synth.make_batch(62, batch_size=30, vocab_size=V)
    words, pos, ner = batch[0].to(DEV), batch[2].to(DEV), batch[3].to(DEV)
    emb = (torch.rand(V, E, device=DEV) + 0.5).requires_grad_()
    pw = (torch.rand(47, 8, device=DEV) + 0.5).requires_grad_()
    nw = (torch.rand(15, 8, device=DEV) + 0.5).requires_grad_()
    rng = torch.tensor([99, 5], dtype=torch.int64, device=DEV)
    x = ops.embed_concat(words, pos, ner, emb, pw, nw, drop_p=0.5, rng_state=rng, subseq=3, topn=150)
    keep = (x != 0)
    frac = keep[words != 0].float().mean().item()
    assert abs(frac - 0.5) < 0.02
    r = torch.randn(x.shape, device=DEV)
    (x * r).sum().backward()
    want = torch.zeros(V, E, device=DEV)
    want.index_put_((words.flatten(),), (r * keep * 2.0)[..., :E].reshape(-1, E), accumulate=True)
    want[0] = 0
    want[150:] = 0                                          # frozen rows (topn)
    assert _rel(emb.grad.cpu(), want.cpu()) < 1e-5
    x2 = ops.embed_concat(words, pos, ner, emb, pw, nw, drop_p=0.5, rng_state=rng + torch.tensor([0, 1], device=DEV),
                          subseq=3)
    assert not torch.equal(x2 != 0, keep)


# ---------------------------------------------------------------- K3 on tcgen05 -------------------------------

@pytest.mark.parametrize('M,N,K', [(128, 16, 32), (50, 200, 360), (2750, 200, 360), (2750, 200, 200), (4096, 512, 360),
                                    (9999, 512, 512), (300, 300, 100), (777, 64, 84), (2750, 200, 400), (640, 72, 600)])
def test_k3_tcgen05_tf32_gemm(M, N, K):
    # tolerance of one TF32 pass (10-bit mantissa operands, fp32 accumulation in TMEM): 2e-3 of the output scale
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    x = torch.randn(M, K, device=DEV, generator=g)
    w = torch.randn(N, K, device=DEV, generator=g)
    dy = torch.randn(M, N, device=DEV, generator=g)
    y = ops.linear_fwd(x, w, 'tf32')
    assert _rel(y.cpu(), (x.double() @ w.double().t()).cpu()) < 2e-3
    dx = ops.linear_dgrad(dy, w, 'tf32')
    assert _rel(dx.cpu(), (dy.double() @ w.double()).cpu()) < 2e-3
    # operands that are exactly representable in TF32 must give the fp32 result up to accumulation order
    xq = (x.view(torch.int32) & -8192).view(torch.float32)
    wq = (w.view(torch.int32) & -8192).view(torch.float32)
    assert _rel(ops.linear_fwd(xq, wq, 'tf32').cpu(), (xq.double() @ wq.double().t()).cpu()) < 1e-5


def test_model_in_tf32_mode_tracks_fp32_mode():
    # stated bf16/tf32-class tolerance for the tensor-core mode: logits within 1e-2 relative of the fp32 path
    torch.manual_seed(3)
    batch = synth.make_batch(9, batch_size=50, vocab_size=800)
    a = GCNTrainer(synth.tacred_opt(vocab_size=800, cuda=True))
    b = GCNTrainer(synth.tacred_opt(vocab_size=800, cuda=True, gemm_mode='tf32'))
    b.model.load_state_dict(a.model.state_dict())
    a.model.eval(); b.model.eval()
    with torch.no_grad():
        la, _ = a.model([t.to(DEV) for t in batch[:-2]])
        lb, _ = b.model([t.to(DEV) for t in batch[:-2]])
    assert _rel(lb.cpu(), la.cpu()) < 1e-2


@pytest.mark.parametrize('M,N,K', [(128, 16, 32), (50, 200, 360), (2750, 200, 360), (2750, 200, 200), (4096, 512, 360),
                                    (9999, 512, 512), (300, 300, 100), (777, 64, 84), (2750, 200, 400), (640, 72, 600)])
def test_k3_tcgen05_3xtf32_gemm_is_fp32_grade(M, N, K):
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    x = torch.randn(M, K, device=DEV, generator=g)
    w = torch.randn(N, K, device=DEV, generator=g)
    dy = torch.randn(M, N, device=DEV, generator=g)
    # the tensor core accumulates with truncation, so the error grows ~linearly with the reduction length
    # (measured 6e-7 at K=32, 3e-6 at K=360); 1e-5 of the output scale is the bound this mode is held to
    assert _rel(ops.linear_fwd(x, w, 'tf32x3').cpu(), (x.double() @ w.double().t()).cpu()) < 1e-5
    assert _rel(ops.linear_dgrad(dy, w, 'tf32x3').cpu(), (dy.double() @ w.double()).cpu()) < 1e-5


def test_tf32_operands_are_truncated_not_rounded():
    # the 3xTF32 split relies on it: the tensor core must see exactly trunc_tf32(a) when handed a raw fp32 a
    x = torch.zeros(128, 32, device=DEV)
    w = torch.zeros(16, 32, device=DEV)
    vals = [1 + 2 ** -11 + 2 ** -12, 1 + 2 ** -10 + 2 ** -11, -(1 + 2 ** -11 + 2 ** -12), 1 + 2 ** -13]
    for i, v in enumerate(vals):
        x[i, 0] = v
    w[0, 0] = 1.0
    y = ops.linear_fwd(x, w, 'tf32')
    want = (x[:4, 0].view(torch.int32) & -8192).view(torch.float32)
    assert torch.equal(y[:4, 0], want)


# ---------------------------------------------------------------- C-GCN encoder without leaving the device -----------

@pytest.mark.gpu
@pytest.mark.parametrize('layers', (1, 2))
def test_device_resident_bilstm_equals_the_packed_sequence_path(layers, monkeypatch):
    """GCN.encode_with_rnn (two unidirectional cuDNN LSTMs over the padded batch, the backward direction on sentences
    reversed inside their own length) == the reference's pack_padded_sequence -> nn.LSTM -> pad_packed_sequence
    (gcn.py:186-197): outputs and every gradient, ragged lengths including length 1 and full width."""
    torch.manual_seed(layers)
    monkeypatch.setattr(torch.backends.cudnn, 'allow_tf32', False)     # cuDNN's LSTM GEMMs in fp32 on both paths
    opt = synth.tacred_opt(vocab_size=300, cuda=True, rnn=True, rnn_hidden=48, rnn_layers=layers, rnn_dropout=0.0)
    model = GCNTrainer(opt).model.gcn_model.gcn
    B, T, D = 9, 23, model.rnn.input_size
    lens = torch.tensor([23, 17, 17, 9, 5, 2, 1, 1, 12])
    masks = (torch.arange(T)[None, :] >= lens[:, None]).cuda()
    x1 = torch.randn(B, T, D, device='cuda', requires_grad=True)
    x2 = x1.detach().clone().requires_grad_()
    g = torch.randn(B, T, 96, device='cuda')
    for training in (False, True):
        model.train(training)
        a = model.encode_with_rnn(x1, masks, B)
        b = model.encode_with_rnn_packed(x2, masks, B)
        assert a.shape == b.shape == (B, T, 96)
        # cuDNN runs different algorithms for packed and padded inputs (observed 8e-6 at |out| <= 1, with or without
        # TF32); the end-to-end bound is the 1e-5 logits parity of cfg3 against the reference (test_model_*)
        assert float((a - b).abs().max()) <= 2e-5
        assert float(a[masks].abs().max()) == 0.0                      # padded positions are zero
    model.zero_grad()
    (a * g).sum().backward()
    ga = {n: p.grad.clone() for n, p in model.rnn.named_parameters()}
    model.zero_grad()
    (b * g).sum().backward()
    for n, p in model.rnn.named_parameters():
        assert float((ga[n] - p.grad).abs().max()) <= 1e-4 * max(1.0, float(p.grad.abs().max())), n
    assert float((x1.grad - x2.grad).abs().max()) <= 1e-4


@pytest.mark.parametrize('n,V,E,Dp,Dn,p,topn', [(70_001, 500, 300, 30, 30, 0.5, 500), (131_072, 3000, 300, 30, 0, 0.0, 2000),
                                                (66_000, 50, 64, 8, 8, 0.3, 50), (65_536, 800, 300, 0, 0, 0.5, 800)])
def test_k5_grouped_backward_equals_the_scatter_backward(n, V, E, Dp, Dn, p, topn):
    """Large batches: rows grouped by word and summed once per word (no fp atomics on the word table) == the scatter
    kernel: same dropout masks (re-derived per row from {seed, step}), unobservable rows skipped, padding row and rows
    >= topn untouched, owner = the word's first live row, gradients ADDED to what the buffers hold."""
    g = torch.Generator(device=DEV).manual_seed(n + V)
    words = torch.randint(0, V, (n,), device=DEV, generator=g)
    pos = torch.randint(0, 47, (n,), device=DEV, generator=g) if Dp else None
    ner = torch.randint(0, 15, (n,), device=DEV, generator=g) if Dn else None
    flags = (torch.rand(n, device=DEV, generator=g) < 0.6).to(torch.uint8)
    dx = torch.randn(n, E + Dp + Dn, device=DEV, generator=g)
    rng = torch.tensor([1234, 7], dtype=torch.int64, device=DEV)
    outs = []
    for min_rows in (1 << 30, 65536):                        # scatter, then grouped
        ops.EMBED_GROUPED_MIN_ROWS = min_rows
        G = torch.full((V, E), 0.25, device=DEV)            # not zero: both paths must accumulate
        gp = torch.zeros(47, Dp, device=DEV) if Dp else None
        gn = torch.zeros(15, Dn, device=DEV) if Dn else None
        owner = torch.full((V,), 0x7fffffff, dtype=torch.int32, device=DEV)
        ops.embed_bwd(dx, flags, words, pos, ner, G, gp, gn, owner, V, E, topn, p, rng, 0xE0)
        torch.cuda.synchronize()
        outs.append((G, gp, gn, owner))
    ops.EMBED_GROUPED_MIN_ROWS = 65536
    (G0, gp0, gn0, o0), (G1, gp1, gn1, o1) = outs
    assert torch.equal(o0, o1)
    assert float((G0 - 0.25).abs().max()) > 0 and torch.equal(G0[0], torch.full((E,), 0.25, device=DEV))
    assert torch.equal(G1[topn:], G0[topn:]) and torch.equal(G1[0], G0[0])
    assert _rel(G1.double().cpu(), G0.double().cpu()) < 2e-5      # ~n/V fp32 additions per entry, in a different order
    if Dp:
        assert _rel(gp1.double().cpu(), gp0.double().cpu()) < 2e-5
    if Dn:
        assert _rel(gn1.double().cpu(), gn0.double().cpu()) < 2e-5


def test_k5_large_batch_forward_and_grouped_backward_share_one_dropout_stream():
    """>= 64 K rows: the resident-grid forward (every thread one 8-column group) and the grouped backward re-derive the
    same Philox mask per (row, column group): d(table) == scatter of (r * keep * scale), keep rate 0.5."""
    V, E, n = 900, 300, 70_003
    g = torch.Generator(device=DEV).manual_seed(9)
    words = torch.randint(0, V, (1, n), device=DEV, generator=g)
    pos = torch.randint(0, 47, (1, n), device=DEV, generator=g)
    ner = torch.randint(0, 15, (1, n), device=DEV, generator=g)
    emb = (torch.rand(V, E, device=DEV, generator=g) + 0.5).requires_grad_()
    pw = (torch.rand(47, 30, device=DEV, generator=g) + 0.5).requires_grad_()
    nw = (torch.rand(15, 30, device=DEV, generator=g) + 0.5).requires_grad_()
    rng = torch.tensor([4321, 11], dtype=torch.int64, device=DEV)
    x = ops.embed_concat(words, pos, ner, emb, pw, nw, drop_p=0.5, rng_state=rng, subseq=0xE0)
    keep = x != 0
    assert abs(keep[0][words[0] != 0].float().mean().item() - 0.5) < 0.01
    scale = ops.drop_scale(0.5)
    full = torch.cat([emb.detach()[words], pw.detach()[pos], nw.detach()[ner]], 2)
    assert torch.equal(x, torch.where(keep, full * scale, torch.zeros_like(full)))
    r = torch.randn(x.shape, device=DEV, generator=g)
    (x * r).sum().backward()
    want = torch.zeros(V, E, device=DEV, dtype=torch.float64)
    want.index_put_((words.flatten(),), (r * keep * scale)[0, :, :E].double(), accumulate=True)
    want[0] = 0
    assert _rel(emb.grad.double().cpu(), want.cpu()) < 1e-5
    wantp = torch.zeros(47, 30, device=DEV, dtype=torch.float64)
    wantp.index_put_((pos.flatten(),), (r * keep * scale)[0, :, E:E + 30].double(), accumulate=True)
    assert _rel(pw.grad.double().cpu(), wantp.cpu()) < 1e-5
