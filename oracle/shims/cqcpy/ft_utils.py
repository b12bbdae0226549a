from kelvin_oracle.cqc import ff, ffv, GP0, uGP0, dGP0, HtoK  # noqa: F401
