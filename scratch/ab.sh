#!/bin/bash
# A/B of two builds of the library on the same box: bench with 24 starts (CUDA-event breakdown inside the library).
set -u
OLD=${OLD:-$PWD/scratch/libmnemcu_r1.so}
for tag in ${TAGS:-old new}; do
  if [ $tag = old ]; then export MNEMCU_LIB=$OLD; else unset MNEMCU_LIB; fi
  python bench.py --starts ${STARTS:-24} --steps 2 --warmup 1 --no-cpu --no-e2e ${EXTRA:-} 2>/dev/null | TAG=$tag python scratch/ab_print.py
done
