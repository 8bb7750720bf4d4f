#!/bin/bash
# One GPU-box visit with everything the evidence index (profiles/README.md) expects:
#   bash gpurun_probe.sh [tag]    ->  gpurun_out/<tag>/ ; then `python tools/summarize_round.py <tag>` copies it into profiles/
bash tools/gpu_round.sh ${1:-r03a}
