#!/bin/bash
for rep in 1 2; do
for f in build/variants/*.so; do
  echo "== $f"
  for n in 8192 3355; do TVL1_SO=$f python scripts/kbench.py iterate $n 2>&1 | tail -2; done
done
done
