"""`python -m esm_b200.extract_cli MODEL FASTA OUT_DIR --include mean per_tok ...` — the bulk extraction driver with the
arguments and output files of the reference's `esm-extract` (/root/reference/scripts/extract.py:15-131): one
`{label}.pt` per sequence holding {"label", "representations": {layer: [len,E]}, "mean_representations",
"bos_representations", "contacts"} as requested by --include.  Differences: the model runs through libesmb200.so
(GPU required, no --nogpu path), the default token budget is larger (the reference's 4096 tokens leave a B200 idle),
and with torchrun the batches are dealt round-robin to the ranks (each rank writes its own files; no collective is
needed because the outputs are files)."""
from __future__ import annotations

import argparse
import os
import pathlib

import torch

from . import pretrained
from .data import FastaBatchedDataset


def create_parser():
    p = argparse.ArgumentParser(description="Extract per-token representations and model outputs for sequences in a FASTA file")
    p.add_argument("model_location", type=str, help="ESM-2 model name (esm2_t33_650M_UR50D, ...) or a local .pt file")
    p.add_argument("fasta_file", type=pathlib.Path)
    p.add_argument("output_dir", type=pathlib.Path)
    p.add_argument("--toks_per_batch", type=int, default=65536, help="maximum batch size in tokens")
    p.add_argument("--repr_layers", type=int, default=[-1], nargs="+")
    p.add_argument("--include", type=str, nargs="+", choices=["mean", "per_tok", "bos", "contacts"], required=True)
    p.add_argument("--truncation_seq_length", type=int, default=1022)
    return p


def run(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    model, alphabet = pretrained.load_model_and_alphabet(args.model_location)
    model = model.eval().to(dev)
    dataset = FastaBatchedDataset.from_file(args.fasta_file)
    batches = dataset.get_batch_indices(args.toks_per_batch, extra_toks_per_seq=1)
    converter = alphabet.get_batch_converter(args.truncation_seq_length)
    args.output_dir.mkdir(parents=True, exist_ok=True)
    return_contacts = "contacts" in args.include
    assert all(-(model.num_layers + 1) <= i <= model.num_layers for i in args.repr_layers)
    repr_layers = [(i + model.num_layers + 1) % (model.num_layers + 1) for i in args.repr_layers]
    n_written = 0
    with torch.no_grad():
        for bi, idxs in enumerate(batches):
            if bi % world != rank:
                continue
            labels, strs, toks = converter([dataset[i] for i in idxs])
            out = model(toks.to(dev, non_blocking=True), repr_layers=repr_layers, return_contacts=return_contacts)
            reps = {layer: t.to("cpu") for layer, t in out["representations"].items()}
            contacts = out["contacts"].to("cpu") if return_contacts else None
            for i, label in enumerate(labels):
                result = {"label": label}
                n = min(args.truncation_seq_length, len(strs[i]))
                # extract.py:104-125: clone() so that the saved file holds only the slice
                if "per_tok" in args.include:
                    result["representations"] = {layer: t[i, 1: n + 1].clone() for layer, t in reps.items()}
                if "mean" in args.include:
                    result["mean_representations"] = {layer: t[i, 1: n + 1].mean(0).clone() for layer, t in reps.items()}
                if "bos" in args.include:
                    result["bos_representations"] = {layer: t[i, 0].clone() for layer, t in reps.items()}
                if return_contacts:
                    result["contacts"] = contacts[i, :n, :n].clone()
                torch.save(result, args.output_dir / f"{label}.pt")
                n_written += 1
    return n_written


def main():
    run(create_parser().parse_args())


if __name__ == "__main__":
    main()
