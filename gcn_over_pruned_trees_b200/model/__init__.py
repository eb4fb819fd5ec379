"""Drop-in for the reference's ``model`` package: with ``gcn_over_pruned_trees_b200/`` on ``PYTHONPATH`` the
reference's ``from model.trainer import GCNTrainer`` (train.py:22, eval.py:13) resolves here."""
