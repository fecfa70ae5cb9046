"""CPU oracle for the Sonic prover hot path — TEST INFRASTRUCTURE ONLY (see oracle/sonic.py)."""
