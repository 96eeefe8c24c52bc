"""CPU oracle (test infrastructure only) -- see oracle/thz_oracle.py."""
