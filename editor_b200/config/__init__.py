from .defaults import _C as cfg  # noqa: F401  (same import surface as the reference: `from config import cfg`)
from .cfgnode import CfgNode  # noqa: F401
