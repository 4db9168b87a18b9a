"""CPU oracles for the MSM hot path.  TEST INFRASTRUCTURE ONLY (see each file's header)."""
