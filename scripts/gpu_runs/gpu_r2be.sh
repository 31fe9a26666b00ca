#!/bin/bash
# round-2 GPU call BE: the hand-over with the express launch (no longer the default) stays covered
timeout 600 python -m pytest tests/test_gpu_machine_scale.py -m gpu -x -q -k "express or 600k" 2>&1 | tail -3
