#!/bin/bash
timeout 300 python tools/debug_shard.py 200003 p2p 2>&1 | grep -v "^\[W\|^W0" | tail -12
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sharded" 2>&1 | tail -5
