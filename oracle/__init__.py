"""CPU oracles of the hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs (cpu_baseline, --impl reference) may import this package;
the product (dav2_b200) never does (tests/test_abi_cpu.py::test_product_never_imports_oracle)."""
