"""protocol_b200 — B200-native task<->worker matching engine for the Prime Protocol
orchestrator's scheduling hot path.  See DESIGN.md and include/prime_match.h."""
from . import abi  # noqa: F401

__all__ = ["abi"]
