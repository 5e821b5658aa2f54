"""Belief-propagation (sum-product) benchmark decoder -- B200 back-end.

Drop-in for ``Continous_LDPC_Decoding/bp_decoder_irreg.py`` of the reference (constructor :23,
``decode_OpenCL_belief_propagation`` :221-286, ``return_errors_all_zero`` :288-293).
"""
from __future__ import annotations

from .. import _lib
from .min_sum_decoder_irreg import _LlrDecoderBase


class BeliefPropagationDecoderClassIrregular(_LlrDecoderBase):
    _algo = _lib.ALGO_BP

    def decode_OpenCL_belief_propagation(self, received_blocks, buffer_in=False, return_buffer=False,
                                         early_termination=None):
        return self._decode_llr(received_blocks, buffer_in, return_buffer, early_termination)

    decode = decode_OpenCL_belief_propagation
