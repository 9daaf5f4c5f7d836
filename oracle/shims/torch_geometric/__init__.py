"""Stand-in for torch-geometric 1.6/1.7 (oracle infrastructure; see oracle/shims/README.md)."""
from . import nn, utils, typing, data, transforms  # noqa: F401
