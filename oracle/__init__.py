"""CPU oracle (test infrastructure). See oracle/ttv_oracle.c. Never imported by nauyaca_b200."""
