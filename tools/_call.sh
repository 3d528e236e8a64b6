python -m pytest tests/test_gpu_golden.py -x -q -k "packed" 2>&1 | tail -5
python tools/sweep_heur_pack.py 16384,65536 0,1,5,6,7,2
