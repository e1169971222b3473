"""MAE reconstruction, the counterpart of ``cinema/examples/inference/mae.py:57-83``: paste the predicted patches of the
masked tokens back into the image grid (every view, not only SAX) and return the mask as an image as well."""

from __future__ import annotations

import torch

from cinema_b200.vit import patchify, unpatchify


@torch.no_grad()
def reconstruct_images(batch: dict[str, torch.Tensor], pred_dict: dict[str, torch.Tensor],
                       enc_mask_dict: dict[str, torch.Tensor], patch_size_dict: dict[str, tuple[int, ...]],
                       grid_size_dict: dict[str, tuple[int, ...]]) -> tuple[dict[str, torch.Tensor], dict[str, torch.Tensor]]:
    """-> ({view: (B, C, *size) image with the masked patches replaced by the predictions}, {view: (B, C, *size) 0 / 1 mask
    image, 1 = reconstructed}).  ``patch_size_dict`` = ``model.dec_patch_size_dict``, ``grid_size_dict[v]`` =
    ``model.enc_down_dict[v].patch_embed.grid_size`` (cinema/examples/inference/mae.py:109-118)."""
    recon, masks = {}, {}
    for view, pred in pred_dict.items():
        image = batch[view].float()
        patches = patchify(image, patch_size_dict[view])
        sel = enc_mask_dict[view]
        patches[sel] = pred.reshape(-1, pred.shape[-1]).to(patches.dtype)
        m = torch.zeros_like(patches)
        m[sel] = 1
        recon[view] = unpatchify(patches, patch_size_dict[view], grid_size_dict[view])
        masks[view] = unpatchify(m, patch_size_dict[view], grid_size_dict[view])
    return recon, masks


@torch.no_grad()
def mae_reconstruct(model, batch: dict[str, torch.Tensor], enc_mask_ratio: float = 0.75):
    """Forward + reconstruction in one call: (loss, reconstructed images, mask images)."""
    model.eval()
    loss, pred_dict, enc_mask_dict, _ = model(batch, enc_mask_ratio=enc_mask_ratio)
    grids = {v: model.enc_down_dict[v].patch_embed.grid_size for v in pred_dict}
    recon, masks = reconstruct_images(batch, pred_dict, enc_mask_dict, model.dec_patch_size_dict, grids)
    return loss, recon, masks
