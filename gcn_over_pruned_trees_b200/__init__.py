"""B200-native (sm_100a) GCN-over-pruned-dependency-trees hot path.

Drop-in for the reference's ``model`` package (``GCNTrainer`` / ``GCNClassifier`` / ``GCNRelationModel`` /
``GCN`` / ``pool`` / ``head_to_tree`` / ``tree_to_adj``): put this directory on ``PYTHONPATH`` and the reference's
``train.py`` / ``eval.py`` resolve ``model.trainer`` to ``gcn_over_pruned_trees_b200/model/trainer.py``.

The device work is hand-written CUDA behind a C-ABI shared library (``csrc/`` -> ``libgptb200.so``, declared in
``include/gpt_b200.h``).  There is no CPU fallback: calling an op without the library raises.
"""
__version__ = '0.1.0'
