"""CPU oracle (test infrastructure only; never imported by gp_b200/)."""
