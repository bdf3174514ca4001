#!/bin/bash
# round 2, call AF: the opt-in cluster decode variant after the shared-code changes of the rework (bit-exactness subset)
BARK_B200_DECODE=cluster timeout -k 3 70 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "generate_tokens or full_size or teacher_forced" 2>&1 | tail -3
