#!/bin/bash
# Sweep of the generated constraint kernel's code-generation / launch settings on one GPU (variants pre-compiled by tools/jit_variants.py):
# chunk:block[:bound] = bytecode instructions per CTA-synchronised chunk : threads per CTA at launch : __launch_bounds__ (1024 -> 64 registers, 512 -> 128).
# Prints the constraint / logup stage times of each setting (NB200_TRACE: the stream is drained at every mark).
for spec in 250:512 125:512 500:512 1000:512 250:256 250:1024 500:1024 250:512:512 500:512:512 250:256:256; do
  IFS=: read c b bound <<< "$spec"
  echo "=== chunk=$c block=$b bound=${bound:-1024}"
  NB200_JIT_BOUND=${bound:-1024} NB200_JIT_CHUNK=$c NB200_JIT_BLOCK=$b NB200_TRACE=1 timeout 120 python bench.py --steps 2 --warmup 1 --no-e2e --no-breakdown --no-cpu-baseline --no-verify 2>&1 >/dev/null \
    | grep -E "constraints: row kernel +1[0-9]\.|constraints: row kernel +[2-9][0-9]\.|constraints: row kernel +[5-9]\.|logup interaction trace +[3-9]\." | tail -4
done
