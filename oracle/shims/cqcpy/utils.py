from kelvin_oracle.cqc import D1, D2, D2u, block_diag  # noqa: F401
