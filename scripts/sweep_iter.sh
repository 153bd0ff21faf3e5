#!/bin/bash
for f in build/variants/*.so; do
  echo "== $f"
  for n in 8192 5243 3355 2684 1638 838; do TVL1_SO=$f python scripts/kbench.py iterate $n 2>&1 | tail -1; done
done
