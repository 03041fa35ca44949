"""CPU oracle (test infrastructure only). See oracle/selective_scan_ref.py."""
