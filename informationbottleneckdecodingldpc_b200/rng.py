"""Philox sub-streams of the on-device random sources (quantizer direct sampling, information bits, AWGN).

Every source draws element i of a call from Philox4x32-10 with key ``key`` and counter ``offset + i``
(``ibldpc_sample_direct`` / ``ibldpc_random_bits`` / ``ibldpc_awgn``); consecutive calls advance ``offset``.
Under a multi-process (one process per GPU) BER run every rank must draw DIFFERENT channel realisations, otherwise
the all-reduced error counters are rank 0's counts times the world size.  The sub-stream is therefore part of the
API: ``set_stream(k)`` selects an independent Philox key derived from ``seed`` and ``k`` (stream 0 keeps ``seed``
itself, so single-process results and the ``ibldpc_uniform`` parity tests are unchanged), and an object whose stream
was never chosen takes the rank of the initialised ``torch.distributed`` process group when it is first used.
"""
from __future__ import annotations

_MASK = (1 << 64) - 1


def stream_key(seed: int, stream: int) -> int:
    """64-bit Philox key of sub-stream ``stream`` of ``seed`` (splitmix64 finaliser over seed + stream * golden ratio).
    ``stream_key(seed, 0) == seed``; different streams give different keys for every seed."""
    seed &= _MASK
    if stream == 0:
        return seed
    z = (seed + (int(stream) & _MASK) * 0x9E3779B97F4A7C15) & _MASK
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
    z ^= z >> 31
    if z == seed:          # cannot collide with stream 0
        z ^= 1
    return z


def default_stream() -> int:
    """Rank of this process in the initialised torch.distributed group, else 0."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return int(dist.get_rank())
    except Exception:
        pass
    return 0


class PhiloxStream:
    """Mixin: ``seed`` / ``_offset`` / ``set_stream``.  Classes call ``self._philox_key()`` when they launch."""

    seed = 20181001
    _offset = 0
    _stream = None      # None = not chosen yet: resolved to the process-group rank at first use

    def set_stream(self, stream: int, offset: int = 0):
        """Select sub-stream ``stream`` (normally the rank) and rewind its counter to ``offset``."""
        self._stream = int(stream)
        self._offset = int(offset)
        return self

    @property
    def stream(self) -> int:
        if self._stream is None:
            self._stream = default_stream()
        return self._stream

    def _philox_key(self) -> int:
        return stream_key(int(self.seed), self.stream)
