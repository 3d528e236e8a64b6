python -m pytest tests/test_gpu_golden.py -x -q -k "packed" 2>&1 | tail -5
FAST=0,1 python tools/sweep_heur_pack.py 1024,16384,65536 0,1
