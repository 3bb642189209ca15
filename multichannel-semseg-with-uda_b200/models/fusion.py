"""Fusion modules of the two-stream (RGB + HHA) MFNet heads (reference models/fusion.py:24-29,53-65).

Only AddFusion is on the MCD hot path (configs "MFNet-AddFusion" / "MFNet-ScoreAddFusion"; note that the
substring test below also routes "ScoreAddFusion" to AddFusion, exactly as the reference does).  The gate /
concat variants are a "next" row of SURVEY.md section 8(f) and raise NotImplementedError for now.
"""
import torch.nn as nn


class AddFusion(nn.Module):
    """x1 + x2.  On low-resolution fp32 score maps this is a 0.2 M-element add; the full-resolution
    ScoreAdd variant never materialises its operands - the head's dual-input upsampling kernel adds them."""

    def forward(self, x1, x2):
        return x1 + x2


def get_fusion_model(fusion_type, n_ch):
    if "GateFusion" in fusion_type:
        raise NotImplementedError("GateFusion is outside the libmcd_sm100 hot-path scope")
    if "AddFusion" in fusion_type:
        return AddFusion()
    if "ConcatFusion" in fusion_type or "ConcatConvFusion" in fusion_type:
        raise NotImplementedError("Concat fusions are outside the libmcd_sm100 hot-path scope")
    raise NotImplementedError()
