"""cinema_b200 -- B200-native (sm_100a) implementation of CineMA's MAE-ViT hot path.

Public surface mirrors the reference package (cinema/__init__.py:3-7,23-34) for the hot path:
``CineMA``, ``ConvViT``, ``patchify`` / ``unpatchify`` and the ViT building blocks.  Importing the package does not
load the CUDA library; the first kernel call does, and fails loudly if it has not been built.
"""

from cinema_b200.convvit import ConvViT
from cinema_b200.mae import CineMA, get_model
from cinema_b200.vit import Attention, Block, PatchEmbed, ViTDecoder, ViTEncoder, get_vit_config, patchify, unpatchify

__all__ = ["Attention", "Block", "CineMA", "ConvViT", "PatchEmbed", "ViTDecoder", "ViTEncoder", "get_model", "get_vit_config",
           "patchify", "unpatchify"]
