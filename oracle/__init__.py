"""Test-only CPU oracle (see aiqmc_oracle.py header).  Never imported by the product path."""
