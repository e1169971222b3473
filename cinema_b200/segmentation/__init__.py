"""Segmentation fine-tuning model of the reference (cinema/segmentation/convunetr.py)."""

from cinema_b200.segmentation.convunetr import ConvUNetR, UpsampleDecoder, check_conv_unetr_enc_dec_compatiblity, get_model
from cinema_b200.segmentation.loss import segmentation_loss

__all__ = ["ConvUNetR", "UpsampleDecoder", "check_conv_unetr_enc_dec_compatiblity", "get_model", "segmentation_loss"]
