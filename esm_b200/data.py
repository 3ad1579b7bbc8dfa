"""FASTA reading and token-budget batching for bulk extraction — host-side mirror of
/root/reference/esm/data.py:19-88 (FastaBatchedDataset).  Pure host string work; nothing here touches the GPU."""
from __future__ import annotations

from typing import List, Sequence, Tuple


class FastaBatchedDataset:
    def __init__(self, sequence_labels: Sequence[str], sequence_strs: Sequence[str]):
        self.sequence_labels = list(sequence_labels)
        self.sequence_strs = list(sequence_strs)

    @classmethod
    def from_file(cls, fasta_file: str) -> "FastaBatchedDataset":
        """data.py:24-54: '>' lines start a record (empty label -> seqnum{line:09d}), other lines are concatenated
        after stripping; duplicate labels are an error."""
        labels: List[str] = []
        seqs: List[str] = []
        cur, buf = None, []
        with open(fasta_file, "r") as fh:
            for line_idx, line in enumerate(fh):
                if line.startswith(">"):
                    if cur is not None:
                        labels.append(cur)
                        seqs.append("".join(buf))
                    name = line[1:].strip()
                    cur = name if name else f"seqnum{line_idx:09d}"
                    buf = []
                else:
                    buf.append(line.strip())
        if cur is not None:
            labels.append(cur)
            seqs.append("".join(buf))
        assert len(set(labels)) == len(labels), "Found duplicate sequence labels"
        return cls(labels, seqs)

    def __len__(self) -> int:
        return len(self.sequence_labels)

    def __getitem__(self, idx: int) -> Tuple[str, str]:
        return self.sequence_labels[idx], self.sequence_strs[idx]

    def get_batch_indices(self, toks_per_batch: int, extra_toks_per_seq: int = 0) -> List[List[int]]:
        """data.py:65-88: sort by length, greedily fill batches so that max_len * batch_size <= toks_per_batch."""
        sizes = sorted((len(s), i) for i, s in enumerate(self.sequence_strs))
        batches: List[List[int]] = []
        buf: List[int] = []
        max_len = 0
        for sz, i in sizes:
            sz += extra_toks_per_seq
            if max(sz, max_len) * (len(buf) + 1) > toks_per_batch and buf:
                batches.append(buf)
                buf, max_len = [], 0
            max_len = max(max_len, sz)
            buf.append(i)
        if buf:
            batches.append(buf)
        return batches
