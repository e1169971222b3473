"""Segmentation fine-tuning model of the reference (cinema/segmentation/convunetr.py)."""

from cinema_b200.segmentation.convunetr import ConvUNetR, UpsampleDecoder, check_conv_unetr_enc_dec_compatiblity, get_model

__all__ = ["ConvUNetR", "UpsampleDecoder", "check_conv_unetr_enc_dec_compatiblity", "get_model"]
