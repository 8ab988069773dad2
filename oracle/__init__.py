"""CPU oracle (test infrastructure only): see oracle/git_oracle.py for the rules on who may import it."""
