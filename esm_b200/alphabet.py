"""Tokeniser / batch converter for the ESM-2 and MSA Transformer paths (host-side mirror of
/root/reference/esm/data.py:91-341).

Only what `esm.pretrained.esm2_*()` / `esm_msa1*()` hand back to a caller is mirrored: the "ESM-1b" alphabet
(33 tokens: <cls>=0 <pad>=1 <eos>=2 <unk>=3, 27 residue/gap symbols, <null_1>, <mask>=32; data.py:151-157 +
constants.py:8-10), the "MSA Transformer" alphabet (same vocabulary, no <eos> appended, data.py:158-164),
`get_batch_converter(truncation_seq_length)`, `BatchConverter.__call__ -> (labels, strs, tokens[int64 B x T])` and
`MSABatchConverter.__call__ -> (labels, strs, tokens[int64 B x R x C])` (data.py:299-341).
Golden token vectors of the reference's tests/test_alphabet.py:17-23,38-44,62-86 are checked in tests/test_alphabet.py.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

# /root/reference/esm/constants.py:8-10
PROTEINSEQ_TOKS = ['L', 'A', 'G', 'V', 'S', 'E', 'R', 'T', 'I', 'D', 'P', 'K', 'Q', 'N', 'F', 'Y', 'M', 'H', 'W', 'C',
                   'X', 'B', 'U', 'Z', 'O', '.', '-']


class Alphabet:
    def __init__(self, standard_toks: Sequence[str], prepend_toks: Sequence[str], append_toks: Sequence[str],
                 prepend_bos: bool = True, append_eos: bool = True, use_msa: bool = False):
        self.standard_toks = list(standard_toks)
        self.prepend_toks = list(prepend_toks)
        self.append_toks = list(append_toks)
        self.prepend_bos = prepend_bos
        self.append_eos = append_eos
        self.use_msa = use_msa
        # data.py:108-112: pad the vocabulary to a multiple of 8 with <null_i> before the appended tokens
        self.all_toks = list(self.prepend_toks) + list(self.standard_toks)
        n_null = (8 - len(self.all_toks) % 8) % 8
        self.all_toks += [f"<null_{i + 1}>" for i in range(n_null)]
        self.all_toks += list(self.append_toks)
        self.tok_to_idx = {tok: i for i, tok in enumerate(self.all_toks)}
        self.unk_idx = self.tok_to_idx["<unk>"]
        self.padding_idx = self.get_idx("<pad>")
        self.cls_idx = self.get_idx("<cls>")
        self.mask_idx = self.get_idx("<mask>")
        self.eos_idx = self.get_idx("<eos>")
        self.all_special_tokens = ["<eos>", "<unk>", "<pad>", "<cls>", "<mask>"]
        self._multi = sorted((t for t in self.all_toks if len(t) > 1), key=len, reverse=True)

    def __len__(self) -> int:
        return len(self.all_toks)

    def get_idx(self, tok: str) -> int:
        return self.tok_to_idx.get(tok, self.unk_idx)

    def get_tok(self, ind: int) -> str:
        return self.all_toks[ind]

    def to_dict(self):
        return dict(self.tok_to_idx)

    @classmethod
    def from_architecture(cls, name: str) -> "Alphabet":
        if name in ("ESM-1b", "roberta_large"):
            return cls(PROTEINSEQ_TOKS, ("<cls>", "<pad>", "<eos>", "<unk>"), ("<mask>",), True, True)
        if name in ("MSA Transformer", "msa_transformer"):
            return cls(PROTEINSEQ_TOKS, ("<cls>", "<pad>", "<eos>", "<unk>"), ("<mask>",), True, False, True)
        raise ValueError(f"esm_b200 covers the ESM-2 and MSA Transformer alphabets only; got {name!r}")

    def tokenize(self, text: str) -> List[str]:
        """Same token stream as data.py:176-247 for the inputs ESM-2 sees: every vocabulary entry is a no-split
        token, whitespace separates, anything else raises KeyError at encode time like the reference."""
        out: List[str] = []
        i, n = 0, len(text)
        while i < n:
            ch = text[i]
            if ch.isspace():
                i += 1
                continue
            if ch == "<":
                for tok in self._multi:
                    if text.startswith(tok, i):
                        out.append(tok)
                        i += len(tok)
                        break
                else:
                    out.append(ch)
                    i += 1
                continue
            out.append(ch)
            i += 1
        return out

    def encode(self, text: str) -> List[int]:
        return [self.tok_to_idx[tok] for tok in self.tokenize(text)]

    def get_batch_converter(self, truncation_seq_length: int = None) -> "BatchConverter":
        if self.use_msa:  # data.py:136-140
            return MSABatchConverter(self, truncation_seq_length)
        return BatchConverter(self, truncation_seq_length)


class BatchConverter:
    """(label, sequence) pairs -> (labels, strs, tokens int64 [B, max_len + 2]) — data.py:253-297."""

    def __init__(self, alphabet: Alphabet, truncation_seq_length: int = None):
        self.alphabet = alphabet
        self.truncation_seq_length = truncation_seq_length

    def __call__(self, raw_batch: Sequence[Tuple[str, str]]):
        a = self.alphabet
        labels, strs = zip(*raw_batch)
        enc = [a.encode(s) for s in strs]
        if self.truncation_seq_length:
            enc = [e[: self.truncation_seq_length] for e in enc]
        max_len = max(len(e) for e in enc)
        bos, eos = int(a.prepend_bos), int(a.append_eos)
        tokens = torch.full((len(enc), max_len + bos + eos), a.padding_idx, dtype=torch.int64)
        for i, e in enumerate(enc):
            if bos:
                tokens[i, 0] = a.cls_idx
            tokens[i, bos: bos + len(e)] = torch.tensor(e, dtype=torch.int64)
            if eos:
                tokens[i, bos + len(e)] = a.eos_idx
        return list(labels), list(strs), tokens


class MSABatchConverter(BatchConverter):
    """One MSA (a list of (label, aligned sequence)) or a list of MSAs -> (labels, strs, tokens int64 [B, R, C]),
    rows and columns padded with <pad> up to the largest MSA of the batch — data.py:299-341."""

    def __call__(self, inputs):
        raw_batch = [inputs] if isinstance(inputs[0][0], str) else inputs
        a = self.alphabet
        max_rows = max(len(msa) for msa in raw_batch)
        max_len = max(len(msa[0][1]) for msa in raw_batch)
        tokens = torch.full((len(raw_batch), max_rows, max_len + int(a.prepend_bos) + int(a.append_eos)),
                            a.padding_idx, dtype=torch.int64)
        labels, strs = [], []
        for i, msa in enumerate(raw_batch):
            if len({len(seq) for _, seq in msa}) != 1:
                raise RuntimeError("Received unaligned sequences for input to MSA, all sequence lengths must be equal.")
            msa_labels, msa_strs, msa_tokens = super().__call__(msa)
            labels.append(msa_labels)
            strs.append(msa_strs)
            tokens[i, : msa_tokens.size(0), : msa_tokens.size(1)] = msa_tokens
        return labels, strs, tokens
