#!/bin/bash
# One GPU, the committed tree: full GPU test-suite, smoke, the default bench line and the reference arm's JSON shape.
mkdir -p gpurun_out
L=gpurun_out/final_sanity.log
{
  echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu -s -k "push_slots" 2>&1 | grep "push info" | head -n 2
  timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 4
  echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -n 2
  echo "== bench"; timeout 900 python bench.py 2>&1 | grep "^{"
} > $L 2>&1
cut -c1-400 $L
