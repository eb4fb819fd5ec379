"""CPU checks of the drop-in test infrastructure: baseline/_ref is a byte-identical copy of the reference, the driver
directory the GPU test runs train.py / eval.py from contains no model/ package, the data fixture is what train.py reads.
The drop-in runs themselves need a GPU (tests/test_gpu_dropin.py)."""
import json
import os
import pickle
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'tools'))
import dropin_run  # noqa: E402
import install_reference  # noqa: E402

needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(dropin_run.REF, 'MANIFEST.sha256')),
                               reason='baseline/_ref not installed (tools/install_reference.py)')


@needs_ref
def test_installed_reference_is_unmodified():
    assert install_reference.verify()
    if os.path.isdir(install_reference.SRC):             # in the build container: against the source tree itself
        want = install_reference.manifest(install_reference.SRC)
        have = install_reference.manifest(install_reference.DST)
        assert want == have
    for f in ('train.py', 'eval.py', 'model/trainer.py', 'model/gcn.py', 'model/tree.py', 'data/loader.py',
              'utils/scorer.py', 'dataset/tacred/train.json'):
        assert os.path.exists(os.path.join(install_reference.DST, f)), f


@needs_ref
@pytest.mark.parametrize('loader', ('reference', 'b200'))
def test_imports_of_the_unmodified_drivers_resolve_to_this_package(loader):
    """train.py:21-24 / eval.py:12-15 import `data.loader`, `model.trainer`, `utils.*`.  Run from the reference tree
    (the script's directory is first on sys.path) with this package on PYTHONPATH: `model` resolves here -- the
    reference's directories are namespace packages, a regular package anywhere on the path wins -- `data` resolves here
    unless GPT_DATA_LOADER=reference, `utils` stays the reference's."""
    import subprocess
    code = ('import sys; sys.path.insert(0, %r)\n'
            'import model.trainer, model.gcn, model.tree, data.loader, utils.scorer, utils.vocab\n'
            'print(model.trainer.__file__); print(data.loader.__file__); print(utils.scorer.__file__)' % dropin_run.REF)
    out = subprocess.run([sys.executable, '-c', code], env=dropin_run._env('b200', loader), cwd=dropin_run.REF,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    trainer, data_loader, scorer = out.stdout.split()[-3:]
    assert trainer == os.path.join(dropin_run.PKG, 'model', 'trainer.py')
    assert data_loader == os.path.join(dropin_run.REF if loader == 'reference' else dropin_run.PKG, 'data', 'loader.py')
    assert scorer == os.path.join(dropin_run.REF, 'utils', 'scorer.py')


@needs_ref
def test_fixture_is_what_train_py_reads(tmp_path):
    data, vocab = dropin_run.make_fixture(str(tmp_path))
    for f in ('train_0.1.json', 'dev.json', 'test.json'):            # train.py:79-84
        assert len(json.load(open(os.path.join(data, f)))) == 20
    words = pickle.load(open(os.path.join(vocab, 'vocab.pkl'), 'rb'))    # utils/vocab.py:62-66
    assert words[:2] == ['<PAD>', '<UNK>'] and len(set(words)) == len(words)
    emb = np.load(os.path.join(vocab, 'embedding.npy'))                   # train.py:147-149
    assert emb.shape == (len(words), 300) and not emb[0].any()
