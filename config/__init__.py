"""Drop-in import path: the reference's callers do ``from config import cfg`` (train_net.py:6, test_net.py:3)."""
from editor_b200.config import cfg, CfgNode  # noqa: F401
