"""B200-native LDPC decoding engine: information-bottleneck (LUT), min-sum and BP flooding
decoders and the AWGN channel quantizer, behind the class API of
mx-strk/InformationBottleneckDecodingLDPC.  Hand-written sm_100a CUDA kernels behind a ctypes
C ABI (include/ibldpc.h); no CPU fallback.
"""
from .AWGN_Channel_Transmission.AWGN_channel import AWGN_channel
from .AWGN_Channel_Transmission.AWGN_Quantizer_BPSK import AWGN_Channel_Quantizer
from .AWGN_Channel_Transmission.LDPC_Transmitter import LDPC_BPSK_Transmitter
from .Continous_LDPC_Decoding.bp_decoder_irreg import BeliefPropagationDecoderClassIrregular
from .Continous_LDPC_Decoding.min_sum_decoder_irreg import Min_Sum_Decoder_class_irregular
from .Discrete_LDPC_decoding.discrete_LDPC_decoder import Discrete_LDPC_Decoder_class
from .Discrete_LDPC_decoding.LDPC_encoder import LDPCEncoder
from .Discrete_LDPC_decoding.discrete_LDPC_decoder_irreg import Discrete_LDPC_Decoder_class_irregular
from .device_array import DeviceArray, pinned_empty

__all__ = ["AWGN_channel", "LDPC_BPSK_Transmitter", "LDPCEncoder", "AWGN_Channel_Quantizer", "BeliefPropagationDecoderClassIrregular", "Min_Sum_Decoder_class_irregular",
           "Discrete_LDPC_Decoder_class", "Discrete_LDPC_Decoder_class_irregular", "DeviceArray", "pinned_empty"]
