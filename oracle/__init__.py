"""CPU oracle (test infrastructure only): see vs_oracle.py."""
