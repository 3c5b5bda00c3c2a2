"""CPU oracle (test infrastructure only -- never imported by immunostruct_b200)."""
