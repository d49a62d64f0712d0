"""CPU oracle of the DQO-MAP hot path — TEST INFRASTRUCTURE ONLY (see oracle/dqo_oracle.c header)."""
