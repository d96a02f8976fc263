"""Measurement scripts that time the oracle / reference arm next to the product (allowed to import oracle/: they live under tests/).  Not collected by pytest."""
