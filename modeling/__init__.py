"""Drop-in import path: the reference's callers do ``from modeling import make_model`` (train_net.py:71, test_net.py:42)."""
from editor_b200.modeling import make_model, build_model, EDITOR  # noqa: F401
