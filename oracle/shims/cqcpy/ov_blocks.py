from kelvin_oracle.cqc import one_e_blocks, two_e_blocks, two_e_blocks_full  # noqa: F401
