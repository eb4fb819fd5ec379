"""Mirror of the reference's ``data`` package for the rows SURVEY.md 8f marks "next": the loader, device-resident.

The reference's packages are namespace packages (no ``__init__.py`` anywhere in its tree), so once this package's parent
directory is on PYTHONPATH the regular packages here -- ``model`` AND ``data`` -- win the import over the reference's own
directories, wherever they sit on sys.path.  ``GPT_DATA_LOADER=reference`` hands ``data.*`` back to the reference's
modules next to the running script (host batches from its own ``data/loader.py``; only ``model`` is replaced)."""
import os
import sys

if os.environ.get('GPT_DATA_LOADER', 'b200') == 'reference' and __name__ == 'data':
    _here = os.path.dirname(os.path.abspath(__file__))
    for _p in sys.path:
        _cand = os.path.join(os.path.abspath(_p or '.'), 'data')
        if _cand != _here and os.path.isfile(os.path.join(_cand, 'loader.py')):
            __path__.insert(0, _cand)
            break
