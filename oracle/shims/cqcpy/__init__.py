"""Shim package re-exporting the restated cqcpy surface (oracle/kelvin_oracle)."""
