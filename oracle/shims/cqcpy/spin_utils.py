from kelvin_oracle.cqc import T1_to_spin, T2_to_spin, D2_to_spin  # noqa: F401
