#!/bin/bash
python bench.py --no-cpu --no-e2e --steps 400 --warmup 100 2>/dev/null | python scripts/summ.py "A"
cp seplib_b200/libsep_m8.so seplib_b200/libsep.so
python bench.py --no-cpu --no-e2e --steps 400 --warmup 100 2>/dev/null | python scripts/summ.py "B"
