"""CPU: FASTA parsing and token-budget batching (mirror of /root/reference/esm/data.py:19-88) against the reference's
behaviour, restated here as properties + a hand-checked example."""
from esm_b200.data import FastaBatchedDataset


def test_fasta_parse_and_batches(tmp_path):
    f = tmp_path / "x.fasta"
    f.write_text(">a desc\nMKT\nVRQ\n>\nAA\n>c\nMKTVRQGMKT\n")
    ds = FastaBatchedDataset.from_file(str(f))
    assert ds.sequence_labels == ["a desc", "seqnum000000003", "c"]
    assert ds.sequence_strs == ["MKTVRQ", "AA", "MKTVRQGMKT"]
    # sorted by length: AA(2) MKTVRQ(6) MKTVRQGMKT(10); +1 token each; budget 14 -> [AA, MKTVRQ] (7*2=14) | [10-mer]
    assert ds.get_batch_indices(14, extra_toks_per_seq=1) == [[1, 0], [2]]
    assert ds.get_batch_indices(1000, extra_toks_per_seq=1) == [[1, 0, 2]]


def test_batches_respect_budget_and_cover_everything():
    import random
    rng = random.Random(0)
    seqs = ["A" * rng.randint(1, 300) for _ in range(200)]
    ds = FastaBatchedDataset([f"s{i}" for i in range(200)], seqs)
    batches = ds.get_batch_indices(1024, extra_toks_per_seq=1)
    seen = sorted(i for b in batches for i in b)
    assert seen == list(range(200))
    for b in batches:
        assert (max(len(seqs[i]) for i in b) + 1) * len(b) <= 1024 or len(b) == 1


def test_extract_file_writer_and_staging_plan(tmp_path):
    """Host-side pieces of the extraction driver: the background writer saves every queued object and reports the
    count; the staging plan covers 256-byte aligned carving."""
    import torch
    from esm_b200.extract_cli import FileWriter, plan_bytes
    w = FileWriter(n_threads=3, depth=4)
    for i in range(25):
        w.put(tmp_path / f"s{i}.pt", {"label": f"s{i}", "x": torch.full((3,), float(i))})
    assert w.close() == 25
    for i in (0, 7, 24):
        r = torch.load(tmp_path / f"s{i}.pt", weights_only=False)
        assert r["label"] == f"s{i}" and float(r["x"][0]) == float(i)
    shapes = [((3, 5, 7), 4), ((3, 7), 4), ((3, 3, 3), 4)]
    need = 0
    for shape, es in shapes:          # what StagingSlot.take() does
        n = es
        for s in shape:
            n *= s
        need = (need + 255) // 256 * 256 + n
    assert plan_bytes(shapes) >= need


def test_extract_file_writer_surfaces_errors(tmp_path):
    import pytest
    import torch
    from esm_b200.extract_cli import FileWriter
    w = FileWriter(n_threads=1)
    w.put(tmp_path / "no_such_dir" / "x.pt", {"x": torch.zeros(1)})
    with pytest.raises(Exception):
        w.close()
