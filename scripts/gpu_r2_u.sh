#!/bin/bash
# round-2 GPU session U (1 GPU): reference arm contract check, PDL A/B incl. the MSA stack
mkdir -p gpurun_out
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-500
echo "== PDL A/B incl. MSA"; timeout 900 python scripts/pdl_ab.py 2>/dev/null | tail -4
