"""Same module path as the reference package ``Discrete_LDPC_decoding`` (IB decoders)."""
