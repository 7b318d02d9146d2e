python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/tune_policy.py 3 "" 2>&1 | tail -1
EVERY=10 python scripts/diag_transient.py 61 "" 2>&1 | tail -7 | cut -c1-110
