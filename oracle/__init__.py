"""CPU oracle of the env-step hot path -- TEST INFRASTRUCTURE ONLY (parity unpinned, see smenv_oracle.c)."""
