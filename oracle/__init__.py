"""CPU oracle for the br hot path — TEST INFRASTRUCTURE ONLY (see oracle/br_oracle.h)."""
