"""Same module path as the reference package ``AWGN_Channel_Transmission`` (channel quantizer)."""
