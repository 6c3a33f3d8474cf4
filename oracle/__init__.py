"""CPU oracle for the b200 worker's hot path — TEST INFRASTRUCTURE ONLY (see oracle/ops.py)."""
