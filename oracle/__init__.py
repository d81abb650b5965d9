"""CPU oracle (test infrastructure only -- see isp_oracle.py header)."""
