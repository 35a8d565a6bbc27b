"""CPU oracle (test infrastructure only; see deepbedmap_oracle.py)."""
