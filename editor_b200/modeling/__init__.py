from .make_model import make_model, build_model, EDITOR  # noqa: F401  (same surface as the reference's modeling/__init__.py:1)
