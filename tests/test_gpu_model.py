"""GPU (-m gpu): the whole path through the reference-facing API (esm_b200.ESM2.forward -> C ABI -> sm_100a kernels)
against (1) committed outputs of the unmodified reference (tests/golden), (2) the CPU oracle on seeded inputs,
(3) size-independent properties at the BASELINE.json model size.

Stated tolerance (fp16 MMA operands, fp32 accumulate / residual / LayerNorm / softmax):
  representations and logits: relative Frobenius error <= 3e-3 / 4e-3 (measured 4e-4 .. 1.3e-3, profiles/r01_parity.txt)
  attention probabilities: max-abs <= 1e-2 (measured <= 2.3e-3);  contacts: max-abs <= 1e-2 (measured <= 8e-4)
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

REL_FRO = 3e-3
REL_FRO_LOGITS = 4e-3
ATT_ABS = 1e-2
CONTACT_ABS = 1e-2


def rel_fro(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def build_model(L, E, H, seed=0):
    from esm_b200 import ESM2
    from oracle.weights import make_state_dict
    sd = make_state_dict(L, E, H, seed=seed)
    model = ESM2(num_layers=L, embed_dim=E, attention_heads=H)
    model.load_state_dict(sd, strict=True)
    return model.eval().cuda(), sd


@pytest.mark.parametrize("name", ["tiny_L2_E128_H2", "mid_L3_E256_H4", "nopad_L2_E128_H2", "t48_15B_like_L2_E256_H2"])
def test_against_reference_golden(name, golden_dir):
    fx = torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)
    cfg = fx["config"]
    model, _ = build_model(cfg["num_layers"], cfg["embed_dim"], cfg["attention_heads"], cfg["seed"])
    out = model(fx["tokens"].cuda(), repr_layers=fx["repr_layers"], need_head_weights=True, return_contacts=True)
    torch.cuda.synchronize()
    assert set(out.keys()) == {"logits", "representations", "attentions", "contacts"}
    for k, ref in fx["representations"].items():
        got = out["representations"][k].cpu()
        assert got.shape == ref.shape
        assert rel_fro(got, ref) <= REL_FRO, (k, rel_fro(got, ref))
    assert rel_fro(out["logits"].cpu(), fx["logits"]) <= REL_FRO_LOGITS
    L, H = cfg["num_layers"], cfg["attention_heads"]
    sub = out["attentions"][:, [0, L - 1]][:, :, [0, H - 1]].cpu()
    assert float((sub - fx["attentions_sub"]).abs().max()) <= ATT_ABS
    if fx["attentions"] is not None:
        assert float((out["attentions"].cpu() - fx["attentions"]).abs().max()) <= ATT_ABS
    assert float((out["contacts"].cpu() - fx["contacts"]).abs().max()) <= CONTACT_ABS


def test_against_oracle_650M_width():
    """4 layers at the 650M width (E=1280, H=20, F=5120), ragged batch, T=300 (3 key blocks, last one partial)."""
    from oracle import esm2_oracle
    from oracle.weights import make_tokens
    L, E, H = 4, 1280, 20
    model, sd = build_model(L, E, H)
    tokens = make_tokens([298, 140, 5], 300, seed=7, n_mask=2)
    ref = esm2_oracle.esm2_forward(sd, L, H, tokens, repr_layers=[0, 2, 4])
    out = model(tokens.cuda(), repr_layers=[0, 2, 4])
    for k in (0, 2, 4):
        r = rel_fro(out["representations"][k].cpu(), ref["representations"][k])
        assert r <= REL_FRO, (k, r)
    assert rel_fro(out["logits"].cpu(), ref["logits"]) <= REL_FRO_LOGITS
    assert "attentions" not in out and "contacts" not in out


def test_layer_level_interface_matches_reference_contract():
    """TransformerLayer.forward(x (T,B,E), self_attn_padding_mask (B,T)) -> (x (T,B,E), attn (H,B,T,T) | None),
    modules.py:120-142."""
    from oracle import esm2_oracle
    model, sd = build_model(1, 128, 2)
    T, B, E = 50, 3, 128
    g = torch.Generator().manual_seed(3)
    x = torch.randn(T, B, E, generator=g)
    pad = torch.zeros(B, T, dtype=torch.bool)
    pad[1, 30:] = True
    ref, probs = esm2_oracle.transformer_layer(x.transpose(0, 1), sd, "layers.0.", 2, pad, True)
    y, attn = model.layers[0](x.cuda(), self_attn_padding_mask=pad.cuda(), need_head_weights=True)
    assert y.shape == (T, B, E) and attn.shape == (2, B, T, T)
    assert rel_fro(y.cpu().transpose(0, 1), ref) <= REL_FRO
    assert float((attn.cpu().transpose(0, 1) - probs).abs().max()) <= ATT_ABS
    y2, attn2 = model.layers[0](x.cuda(), self_attn_padding_mask=pad.cuda())
    assert attn2 is None and torch.equal(y2, y)


def test_full_size_properties_650M():
    """BASELINE.json configs[1] model (33 x 1280 x 20 heads) at L=1024: properties that need no CPU oracle run.
    (a) a sequence embedded alone equals the same sequence embedded inside a ragged batch, bit for bit
        (sequences are independent, esm2.py:77-144; padding keys get exactly zero probability);
    (b) batch order does not matter; (c) outputs are finite and post-LayerNorm statistics are sane."""
    from oracle.weights import make_tokens
    torch.manual_seed(0)
    from esm_b200 import pretrained
    model, alphabet = pretrained.esm2_t33_650M_UR50D(allow_random_init=True)
    model = model.cuda()
    tokens = make_tokens([1022, 700, 1022, 333], 1024, seed=11).cuda()
    out = model(tokens, repr_layers=[33])["representations"][33]
    assert out.shape == (4, 1024, 1280) and bool(torch.isfinite(out).all())
    perm = torch.tensor([2, 0, 3, 1], device="cuda")
    out_p = model(tokens[perm], repr_layers=[33])["representations"][33]
    assert torch.equal(out_p, out[perm])
    alone = model(tokens[1:2, :702], repr_layers=[33])["representations"][33]
    assert torch.equal(alone[0], out[1, :702])
    row_mean = out[0].mean(-1).abs().max()
    assert float(row_mean) < 1.0


def test_3B_width_contacts():
    """BASELINE.json configs[3] width (E=2560, H=40, F=10240; 2 of the 36 layers), need_head_weights / contacts path,
    ragged batch, T=260 (3 key blocks of 128 for the probability kernel, 5 blocks of 64 for the forward kernel)."""
    from oracle import esm2_oracle
    from oracle.weights import make_tokens
    L, E, H = 2, 2560, 40
    model, sd = build_model(L, E, H)
    tokens = make_tokens([258, 100], 260, seed=3, n_mask=1)
    ref = esm2_oracle.esm2_forward(sd, L, H, tokens, repr_layers=[1, 2], return_contacts=True)
    out = model(tokens.cuda(), repr_layers=[1, 2], return_contacts=True)
    for k in (1, 2):
        assert rel_fro(out["representations"][k].cpu(), ref["representations"][k]) <= REL_FRO
    assert rel_fro(out["logits"].cpu(), ref["logits"]) <= REL_FRO_LOGITS
    assert out["attentions"].shape == (2, L, H, 260, 260)
    assert float((out["attentions"].cpu() - ref["attentions"]).abs().max()) <= ATT_ABS
    assert float((out["contacts"].cpu() - ref["contacts"]).abs().max()) <= CONTACT_ABS


def test_bulk_embedder_host_to_host():
    """esm_b200.extract.BulkEmbedder (the e2e call bench.py times): pinned host tokens in, host mean / bos / per-token
    representations out, micro-batched with copy/compute overlap — must equal one direct forward."""
    from esm_b200.extract import BulkEmbedder
    from oracle.weights import make_tokens
    model, _ = build_model(2, 128, 2)
    tokens = make_tokens([60, 33, 47, 12, 60, 5, 29], 62, seed=9)
    direct = model(tokens.cuda(), repr_layers=[2])["representations"][2].cpu()
    emb = BulkEmbedder(model, include=("mean", "bos", "per_tok"), micro_batch=3)
    res = emb.embed(tokens.pin_memory())
    assert torch.equal(res["per_tok"], direct)
    assert torch.equal(res["bos"], direct[:, 0])
    lengths = [60, 33, 47, 12, 60, 5, 29]
    for i, n in enumerate(lengths):
        torch.testing.assert_close(res["mean"][i], direct[i, 1:n + 1].mean(0), atol=1e-5, rtol=1e-5)
    assert emb.d2h_bytes == res["per_tok"].numel() * 4 + 2 * res["mean"].numel() * 4


def test_extract_cli_writes_reference_schema(tmp_path):
    """python -m esm_b200.extract_cli: files and keys of scripts/extract.py:104-131, values against the oracle."""
    import argparse
    from esm_b200 import extract_cli, pretrained
    from oracle import esm2_oracle
    from oracle.weights import make_state_dict
    from esm_b200 import ESM2
    L, E, H = 2, 128, 2
    sd = make_state_dict(L, E, H)
    ckpt = tmp_path / "esm2_tiny.pt"
    torch.save({"cfg": {"model": {"encoder_layers": L, "encoder_embed_dim": E, "encoder_attention_heads": H,
                                  "token_dropout": True}},
                "model": {("encoder.sentence_encoder." + k): v for k, v in sd.items()}}, ckpt)
    fasta = tmp_path / "in.fasta"
    seqs = {"p1": "MKTVRQERLKSIVRILERSKEPVSGAQ", "p2": "KALTARQQEVFDLIRD", "p3": "MKT"}
    fasta.write_text("".join(f">{k}\n{v}\n" for k, v in seqs.items()))
    outdir = tmp_path / "out"
    args = argparse.Namespace(model_location=str(ckpt), fasta_file=fasta, output_dir=outdir, toks_per_batch=64,
                              repr_layers=[-1], include=["mean", "per_tok", "bos", "contacts"],
                              truncation_seq_length=1022)
    n = extract_cli.run(args)
    assert n == 3
    model, alphabet = pretrained.load_model_and_alphabet(str(ckpt))
    for label, seq in seqs.items():
        r = torch.load(outdir / f"{label}.pt", weights_only=False)
        assert set(r.keys()) == {"label", "representations", "mean_representations", "bos_representations", "contacts"}
        _, _, tok = alphabet.get_batch_converter()([(label, seq)])
        ref = esm2_oracle.esm2_forward(sd, L, H, tok, repr_layers=[L], return_contacts=True)
        want = ref["representations"][L][0, 1:len(seq) + 1]
        assert r["representations"][L].shape == (len(seq), E)
        assert rel_fro(r["representations"][L], want) <= REL_FRO
        assert rel_fro(r["mean_representations"][L], want.mean(0)) <= REL_FRO
        assert r["contacts"].shape == (len(seq), len(seq))
        assert float((r["contacts"] - ref["contacts"][0]).abs().max()) <= CONTACT_ABS


@pytest.mark.parametrize("T,eos", [(40, True), (130, True), (600, True), (77, False)])
def test_contact_head_native_accumulation_matches_torch_formula(T, eos):
    """esmb200_contact_accumulate (one pass over each layer's maps) + the small [B,S,S] tail against the same formula
    evaluated with PyTorch ops on the same CUDA tensors; both row-tile variants (S <= 512 / <= 1024), eos masking with
    padded sequences, and the MSA case (no eos appended)."""
    from esm_b200.model import ContactPredictionHead
    torch.manual_seed(T)
    B, L, H = 3, 2, 5
    head = ContactPredictionHead(L * H, True, eos, eos_idx=2).cuda()
    with torch.no_grad():
        head.regression.weight.normal_(0, 1.5)
        head.regression.bias.fill_(0.3)
    tok = torch.randint(4, 24, (B, T))
    tok[:, 0] = 0
    if eos:
        tok[0, -1] = 2
        tok[1, T - 7] = 2
        tok[1, T - 6:] = 1
        tok[2, T // 2] = 2
        tok[2, T // 2 + 1:] = 1
    att = torch.rand(B, L, H, T, T).softmax(-1).cuda()
    tok = tok.cuda()
    with torch.no_grad():
        got = head(tok, att)
        lo, hi = 1, (T - 1 if eos else T)
        ref = head._forward_torch(tok, att, head.regression.weight.view(L, H), lo, hi)
    assert got.shape == (B, hi - lo, hi - lo)
    assert float((got - ref).abs().max()) <= 2e-5


def test_contacts_are_bit_reproducible():
    """the contact head used float atomics across CTAs; now every sum has a fixed order."""
    from oracle.weights import make_tokens
    model, _ = build_model(3, 256, 4)
    tokens = make_tokens([300, 211, 40], 302, seed=2).cuda()
    outs = [model(tokens, return_contacts=True) for _ in range(4)]
    torch.cuda.synchronize()
    for o in outs[1:]:
        assert torch.equal(o["contacts"], outs[0]["contacts"])
        assert torch.equal(o["attentions"], outs[0]["attentions"])
        assert torch.equal(o["logits"], outs[0]["logits"])

