"""CPU: known-answer token vectors from the reference's own tests (/root/reference/tests/test_alphabet.py:17-23,
38-44, 62-86) — the only offline golden vectors the reference holds for this path's inputs."""
import torch

from esm_b200.alphabet import Alphabet


def test_vocabulary_layout():
    a = Alphabet.from_architecture("ESM-1b")
    assert len(a) == 33
    assert (a.cls_idx, a.padding_idx, a.eos_idx, a.unk_idx, a.mask_idx) == (0, 1, 2, 3, 32)
    assert a.get_idx("L") == 4 and a.get_idx("C") == 23 and a.get_tok(31) == "<null_1>"
    assert a.prepend_bos and a.append_eos


def test_esm1b_golden_tokens():
    a = Alphabet.from_architecture("ESM-1b")
    data = [("protein1", "MKTVRQG"), ("protein2 with mask", "KALTA<mask>ISQP"), ("protein3", "K A <mask> I S Q")]
    labels, strs, toks = a.get_batch_converter()(data)
    expected = torch.tensor([
        [0, 20, 15, 11, 7, 10, 16, 6, 2, 1, 1, 1],
        [0, 15, 5, 4, 11, 5, 32, 12, 8, 16, 14, 2],
        [0, 15, 5, 32, 12, 8, 16, 2, 1, 1, 1, 1],
    ])
    assert torch.equal(toks, expected)
    assert labels == ["protein1", "protein2 with mask", "protein3"] and strs[0] == "MKTVRQG"


def test_esm1b_golden_tokens_truncation():
    a = Alphabet.from_architecture("ESM-1b")
    data = [("protein1", "MKTVRQGMKTVRQG"), ("protein2 with mask", "KALTA<mask>ISQPISQP"),
            ("protein3", "K A <mask> I S Q")]
    _, _, toks = a.get_batch_converter(truncation_seq_length=10)(data)
    expected = torch.tensor([
        [0, 20, 15, 11, 7, 10, 16, 6, 20, 15, 11, 2],
        [0, 15, 5, 4, 11, 5, 32, 12, 8, 16, 14, 2],
        [0, 15, 5, 32, 12, 8, 16, 2, 1, 1, 1, 1],
    ])
    assert torch.equal(toks, expected)


def test_unknown_symbol_raises_like_reference():
    a = Alphabet.from_architecture("ESM-1b")
    try:
        a.encode("MKJ")
    except KeyError:
        return
    raise AssertionError("expected KeyError for a symbol outside the vocabulary")


def test_msa_transformer_golden_tokens():
    """/root/reference/tests/test_alphabet.py:62-86 (esm_msa1b alphabet: <cls> prepended, no <eos>, tokens [1, R, C])."""
    a = Alphabet.from_architecture("MSA Transformer")
    assert a.prepend_bos and not a.append_eos and a.use_msa and len(a) == 33
    data = [("protein1", "MKTVRQG"), ("protein2", "KALTRAI"), ("protein3", "KAAISQQ")]
    labels, strs, toks = a.get_batch_converter()(data)
    expected = torch.tensor([[[0, 20, 15, 11, 7, 10, 16, 6], [0, 15, 5, 4, 11, 10, 5, 12], [0, 15, 5, 5, 12, 8, 16, 16]]])
    assert torch.equal(toks, expected)
    assert labels == [["protein1", "protein2", "protein3"]]


def test_msa_batch_converter_pads_rows_and_columns():
    a = Alphabet.from_architecture("msa_transformer")
    batch = [[("a", "MKT"), ("b", "MRT")], [("c", "MKTVR")]]
    _, _, toks = a.get_batch_converter()(batch)
    assert toks.shape == (2, 2, 6)
    assert toks[0, 0].tolist() == [0, 20, 15, 11, 1, 1] and toks[1, 1].tolist() == [1] * 6
    try:
        a.get_batch_converter()([("a", "MKT"), ("b", "MK")])
    except RuntimeError:
        return
    raise AssertionError("unaligned MSA must raise like the reference")
