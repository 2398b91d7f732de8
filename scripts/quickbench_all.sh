#!/bin/bash
# strip/lines kernel times at 2048^2 for every kernel variant (run on the GPU box)
cd "$(dirname "$0")/.."
for rheo in mevp bbm; do
  QB_RHEO=$rheo python scripts/quickbench.py
  QB_RHEO=$rheo QB_DISTORT=1 python scripts/quickbench.py
  QB_RHEO=$rheo QB_SPH=1 python scripts/quickbench.py
done
