"""CPU oracle (test infrastructure only -- see adainvc_oracle.py header)."""
