"""Information-bottleneck decoder for irregular LDPC codes with message alignment -- B200 back-end.

Drop-in for ``Discrete_LDPC_decoding/discrete_LDPC_decoder_irreg.py`` of the reference.
"""
from __future__ import annotations

from .discrete_LDPC_decoder import Discrete_LDPC_Decoder_class


class Discrete_LDPC_Decoder_class_irregular(Discrete_LDPC_Decoder_class):
    """Reference signature: discrete_LDPC_decoder_irreg.py:34-41 (``match`` is the string
    'true' / 'false' the reference pastes into ``#define MATCH``)."""

    _irregular = True

    def __init__(self, filename, imax_, cardinality_T_channel_, cardinality_T_decoder_ops_,
                 Trellis_checknode_vector_a_, Trellis_varnode_vector_a_,
                 matching_vector_checknode_, matching_vector_varnode_, msg_at_time_, match='true'):
        self._load_graph(filename)
        self.imax = imax_
        self.cardinality_T_channel = cardinality_T_channel_
        self.cardinality_T_decoder_ops = cardinality_T_decoder_ops_
        self.update_trellis_vectors(Trellis_checknode_vector_a_, Trellis_varnode_vector_a_)
        self._set_rate()                       # R_c, data_len (discrete_LDPC_decoder_irreg.py:69-100)
        self.msg_at_time = msg_at_time_
        self.matching_vector_checknode = matching_vector_checknode_
        self.matching_vector_varnode = matching_vector_varnode_
        self.match = match if isinstance(match, str) else ('true' if match else 'false')
        self._post_init()

    def set_code_parameters(self):
        self._set_rate()

    def return_errors_all_zero(self, varnode_output_buffer):
        """Only the first data_len = int(R_c*N) rows count (discrete_LDPC_decoder_irreg.py:343-349)."""
        return self._errors(varnode_output_buffer, int(self.data_len))
