"""TEST INFRASTRUCTURE ONLY."""
from . import _GateInst as Gate  # noqa: F401
