"""Test-only CPU oracle (see oracle/model.py header)."""
