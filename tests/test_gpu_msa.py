"""GPU (-m gpu): the MSA axial block (esm_b200.msa.AxialTransformerLayer -> C ABI -> sm_100a kernels) against the
committed outputs of the reference's AxialTransformerLayer and against the CPU oracle. Same tolerance as the ESM-2
path (fp16 operands): rel-Frobenius <= 3e-3 on the layer output, probabilities max-abs <= 1e-2."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_fro(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def build(E, Fd, H, seed=0):
    from esm_b200.msa import AxialTransformerLayer
    from oracle.msa_oracle import make_axial_state_dict
    sd = make_axial_state_dict(E, Fd, seed=seed)
    layer = AxialTransformerLayer(E, Fd, H)
    layer.load_state_dict({k[len("layers.0."):]: v for k, v in sd.items()}, strict=True)
    return layer.eval().cuda(), sd


def test_axial_layer_against_reference_golden(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "msa_mid_E256_H4.pt"), weights_only=False)
    cfg = fx["config"]
    g = torch.Generator().manual_seed(cfg["x_seed"])
    x = torch.randn(cfg["B"], cfg["R"], cfg["C"], cfg["E"], generator=g)
    layer, _ = build(cfg["E"], cfg["F"], cfg["H"])
    out, col, row = layer(x.permute(1, 2, 0, 3).cuda(), need_head_weights=True)  # reference layout (R,C,B,E)
    assert out.shape == (cfg["R"], cfg["C"], cfg["B"], cfg["E"])
    assert rel_fro(out.permute(2, 0, 1, 3).cpu(), fx["out"]) <= 3e-3
    assert row.shape == fx["row_attn"].shape
    assert float((row.cpu() - fx["row_attn"]).abs().max()) <= 1e-2
    assert float((col[:, :4].cpu() - fx["col_attn_sample"]).abs().max()) <= 1e-2


def test_axial_layer_against_oracle_msa1b_width():
    """esm_msa1b width (E=768, H=12, F=3072), 2 MSAs of 16 rows x 192 columns."""
    from oracle import msa_oracle
    layer, sd = build(768, 3072, 12)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 16, 192, 768, generator=g)
    ref = msa_oracle.axial_layer(x, sd, "layers.0.", 12)
    out = layer(x.permute(1, 2, 0, 3).cuda())
    assert rel_fro(out.permute(2, 0, 1, 3).cpu(), ref) <= 3e-3


def test_padded_msa_is_refused_not_silently_wrong():
    layer, _ = build(128, 512, 2)
    x = torch.randn(4, 10, 1, 128).cuda()
    mask = torch.zeros(1, 4, 10, dtype=torch.bool).cuda()
    mask[0, :, -1] = True
    with pytest.raises(NotImplementedError):
        layer(x, self_attn_padding_mask=mask)
