"""Same module path as the reference package ``Continous_LDPC_Decoding`` (min-sum / BP)."""
