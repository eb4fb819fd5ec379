"""The reference's UNMODIFIED drivers on this package, and checkpoints exchanged both ways.

baseline/_ref/ is a byte-identical copy of the reference (tools/install_reference.py; git-ignored, travels to the GPU
box with the snapshot).  train.py / eval.py are executed as subprocesses from a driver directory that has no model/, so
`from model.trainer import GCNTrainer` (train.py:22, eval.py:13) resolves to gcn_over_pruned_trees_b200/model through
PYTHONPATH -- argparse's opt (topn=1e10 as a float, the deprel_* keys, cuda from torch.cuda.is_available()),
helper.save_config, checkpoint_epoch_N.pt -> best_model.pt (train.py:329-337) and eval.py's load_config -> GCNTrainer(opt)
-> load are all the reference's own code.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(REPO, 'tools'))
import dropin_run  # noqa: E402

needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(dropin_run.REF, 'train.py')),
                               reason='baseline/_ref not installed (tools/install_reference.py)')

REFERENCE_KEYS = ['gcn_model.emb.weight', 'gcn_model.pos_emb.weight', 'gcn_model.ner_emb.weight',
                  'gcn_model.deprel_emb.weight', 'gcn_model.gcn.emb.weight', 'gcn_model.gcn.pos_emb.weight',
                  'gcn_model.gcn.ner_emb.weight', 'gcn_model.gcn.deprel_emb.weight', 'gcn_model.gcn.W.0.weight',
                  'gcn_model.gcn.W.0.bias', 'gcn_model.gcn.W.1.weight', 'gcn_model.gcn.W.1.bias',
                  'gcn_model.out_mlp.0.weight', 'gcn_model.out_mlp.0.bias', 'gcn_model.out_mlp.2.weight',
                  'gcn_model.out_mlp.2.bias', 'classifier.weight', 'classifier.bias']


@needs_ref
@pytest.mark.parametrize('loader', ('reference', 'b200'))
def test_unmodified_train_py_and_eval_py_run_on_this_package(tmp_path, loader):
    """SURVEY.md section 4's two commands.  loader='reference': only `model` is replaced (host batches from the
    reference's data/loader.py); loader='b200': `data` resolves here too (device-resident batches, K9)."""
    work = str(tmp_path)
    out, model_dir = dropin_run.run_train(work, model='b200', loader=loader, epochs=2)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert 'Finetune all embeddings.' in out.stdout             # this package's model was constructed ...
    assert 'cuda : True' in out.stdout                          # ... on the GPU, by argparse's default
    assert 'epoch 2: train_loss' in out.stdout
    losses = [float(line.split('loss = ')[1].split(' ')[0]) for line in out.stdout.splitlines() if ', loss = ' in line]
    assert len(losses) == 2 and all(np.isfinite(losses))
    for f in ('best_model.pt', 'config.json', 'vocab.pkl', 'logs.txt'):
        assert os.path.exists(os.path.join(model_dir, f)), f
    cfg = json.load(open(os.path.join(model_dir, 'config.json')))
    assert cfg['topn'] == 1e10 and cfg['cuda'] is True and cfg['num_class'] == 42
    ckpt = torch.load(os.path.join(model_dir, 'best_model.pt'), map_location='cpu')
    assert sorted(ckpt['model'].keys()) == sorted(REFERENCE_KEYS)
    assert ckpt['model']['gcn_model.gcn.W.0.weight'].shape == (200, 360)
    ev = dropin_run.run_eval(work, model_dir, model='b200', loader=loader)
    assert ev.returncode == 0, ev.stdout[-3000:] + ev.stderr[-3000:]
    assert 'Evaluation ended.' in ev.stdout and 'test set evaluate result' in ev.stdout
    # the library that ran is this package's (the subprocess would have died on a missing .so: no CPU fallback)
    assert 'Traceback' not in ev.stderr


def _ref_side(mode, work, ckpt, out, opt=None):
    cmd = [sys.executable, os.path.join(HERE, 'ref_side.py'), mode, '--workdir', work, '--ckpt', ckpt, '--out', out]
    if opt is not None:
        path = os.path.join(work, 'opt.json')
        json.dump(opt, open(path, 'w'))
        cmd += ['--opt', path]
    env = dict(os.environ, PYTHONPATH='')
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=work)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def _predict_here(work, ckpt):
    """eval.py:40-66 on this package: load_config -> GCNTrainer(opt) -> load -> predict over the reference's batches."""
    from gcn_over_pruned_trees_b200 import torch_utils
    from gcn_over_pruned_trees_b200.data.loader import DataLoader
    from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer
    import pickle

    class V(object):
        pass
    vocab = V()
    vocab.id2word = pickle.load(open(os.path.join(work, 'vocab', 'vocab.pkl'), 'rb'))
    vocab.word2id = {w: i for i, w in enumerate(vocab.id2word)}
    vocab.size = len(vocab.id2word)
    opt = torch_utils.load_config(ckpt)
    opt['cuda'], opt['cpu'] = True, False
    trainer = GCNTrainer(opt)
    trainer.load(ckpt)
    trainer.opt['cuda'], trainer.opt['cpu'] = True, False
    out = {}
    for split in ('dev', 'test'):
        loader = DataLoader(os.path.join(work, 'data', split + '.json'), 50, trainer.opt, vocab, evaluation=True)
        preds, probs, losses = [], [], []
        for b in loader:
            p, q, loss = trainer.predict(b)
            preds += p
            probs += q
            losses.append(loss)
        out[split + '_preds'], out[split + '_probs'], out[split + '_loss'] = (np.array(preds), np.array(probs),
                                                                               np.array(losses))
    return trainer, out


def _assert_same_predictions(a, b):
    for split in ('dev', 'test'):
        assert np.array_equal(a[split + '_preds'], b[split + '_preds'])
        ref = b[split + '_probs']
        assert np.abs(a[split + '_probs'] - ref).max() <= 1e-5 * np.abs(ref).max()
        assert np.allclose(a[split + '_loss'], b[split + '_loss'], rtol=1e-5, atol=0)


REF_OPT = dict(emb_dim=300, ner_dim=30, pos_dim=30, hidden_dim=200, num_layers=2, input_dropout=0.5, gcn_dropout=0.5,
               word_dropout=0.04, topn=1e10, lower=False, prune_k=1, conv_l2=0, pooling='max', pooling_l2=0.003,
               mlp_layers=2, no_adj=False, rnn=False, rnn_hidden=200, rnn_layers=1, rnn_dropout=0.5, lr=0.3,
               optim='sgd', max_grad_norm=5.0, adj_type='regular', deprel_emb_dim=200, deprel_dropout=0.5,
               deprel_self_loop=True, deprel_directed=False, use_bert_embeddings=False, emb_dropout=0.0,
               dataset='tacred', deprel_attn=False, deprel_alpha=1.0, edge_keep_prob=1.0, deprel_keep_prop=1.0,
               deprel_max_depth=2, batch_size=50)


@needs_ref
@pytest.mark.parametrize('rnn', (False, True))
def test_checkpoint_saved_by_the_reference_loads_here_with_equal_predict(tmp_path, rnn):
    work = str(tmp_path)
    dropin_run.make_fixture(work)
    ckpt, ref_out = os.path.join(work, 'ref_model.pt'), os.path.join(work, 'ref_predict.npz')
    _ref_side('save', work, ckpt, ref_out, dict(REF_OPT, rnn=rnn))     # reference: train 3 steps on CPU, save, predict
    _, here = _predict_here(work, ckpt)
    _assert_same_predictions(here, np.load(ref_out))


@needs_ref
def test_checkpoint_saved_here_loads_in_the_reference_with_equal_predict(tmp_path):
    from gcn_over_pruned_trees_b200 import synth
    from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer
    work = str(tmp_path)
    dropin_run.make_fixture(work)
    import pickle
    V = len(pickle.load(open(os.path.join(work, 'vocab', 'vocab.pkl'), 'rb')))
    torch.manual_seed(5)
    opt = dict(REF_OPT, vocab_size=V, num_class=42, cuda=True, cpu=False)
    trainer = GCNTrainer(opt)
    trainer.model.train()
    for i in range(3):                                          # move the weights with this package's own step
        trainer.train_step(synth.make_batch(40 + i, batch_size=20, vocab_size=V))
    torch.cuda.synchronize()
    ckpt, ref_out = os.path.join(work, 'b200_model.pt'), os.path.join(work, 'ref_predict.npz')
    trainer.save(ckpt, 1)
    _ref_side('load', work, ckpt, ref_out)                      # the reference's load_config -> GCNTrainer -> load
    _, here = _predict_here(work, ckpt)
    _assert_same_predictions(here, np.load(ref_out))
