#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu --no-e2e --steps 500 --warmup 100 2>/dev/null | python scripts/summ.py "default"
