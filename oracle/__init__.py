"""CPU oracle — test infrastructure only (see pm_oracle.h)."""
