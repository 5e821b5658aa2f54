"""CPU oracle of the LDPC decoding hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm may import
anything from this package; the product package never does (see oracle/README.md)."""
