#!/usr/bin/env python
"""Place the UNMODIFIED reference under baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box).

The reference is a plain script tree without setup.py / pyproject.toml, so `pip install --target baseline/_ref
/root/reference` has nothing to build ("neither 'setup.py' nor 'pyproject.toml' found"); what pip would have produced
for a pure-Python tree -- a byte-for-byte copy of the sources -- is made directly.  Nothing is edited: every file's
sha256 is written to baseline/_ref/MANIFEST.sha256 and re-checked by tests/test_dropin_cpu.py.

Used by: bench.py --impl reference / cpu_baseline (kind "reference": the reference's own GCNTrainer on the host cores),
tests/test_gpu_dropin.py (train.py / eval.py unmodified, `model` resolved to this package), tests/golden generators.
Product code never imports it.
"""
import hashlib
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get('GPT_REFERENCE_SRC', '/root/reference')
DST = os.path.join(REPO, 'baseline', '_ref')
SKIP_DIRS = {'.git', 'fig', '__pycache__'}


def _files(root):
    for base, dirs, files in os.walk(root):
        dirs[:] = sorted(d for d in dirs if d not in SKIP_DIRS)
        for f in sorted(files):
            if f.endswith('.pyc') or f == 'MANIFEST.sha256':
                continue
            yield os.path.relpath(os.path.join(base, f), root)


def _sha(path):
    with open(path, 'rb') as f:
        return hashlib.sha256(f.read()).hexdigest()


def manifest(root):
    return {rel: _sha(os.path.join(root, rel)) for rel in _files(root)}


def install(force=False):
    """Copy SRC -> DST when SRC exists; returns DST, or None when there is neither a source nor an installed copy."""
    if not os.path.isdir(SRC):
        return DST if os.path.exists(os.path.join(DST, 'MANIFEST.sha256')) else None
    want = manifest(SRC)
    have_path = os.path.join(DST, 'MANIFEST.sha256')
    if not force and os.path.exists(have_path):
        have = dict(line.split()[::-1] for line in open(have_path).read().splitlines() if line.strip())
        if have == want and all(os.path.exists(os.path.join(DST, r)) for r in want):
            return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for rel in want:
        out = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), out)
        os.chmod(out, 0o644)
    with open(have_path, 'w') as f:
        for rel, digest in sorted(want.items()):
            f.write('%s  %s\n' % (digest, rel))
    return DST


def verify():
    """True when every installed file still has the digest recorded at install time."""
    path = os.path.join(DST, 'MANIFEST.sha256')
    if not os.path.exists(path):
        return False
    for line in open(path).read().splitlines():
        digest, rel = line.split(None, 1)
        p = os.path.join(DST, rel.strip())
        if not os.path.exists(p) or _sha(p) != digest:
            return False
    return True


if __name__ == '__main__':
    out = install(force='--force' in sys.argv)
    print(out if out else 'reference source %s not present and no installed copy' % SRC)
