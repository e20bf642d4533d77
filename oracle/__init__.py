"""CPU oracle — test infrastructure only (see ms_hgnn_oracle.py)."""
