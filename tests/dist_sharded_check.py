#!/usr/bin/env python
"""torchrun entry point of the sharded-vs-single-GPU parity check (the logic lives in tools/sharded_check.py, which
bench.py --gpus N > 1 also runs before timing)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import sharded_check

if __name__ == "__main__":
    sharded_check.main()
