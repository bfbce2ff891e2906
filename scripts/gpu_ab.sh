timeout 60 python -m pytest tests/test_fnd_loop.py -x -q -m gpu -k "reference_functions" 2>&1 | tail -15
