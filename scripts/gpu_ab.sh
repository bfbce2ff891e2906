timeout 25 python scripts/pow_dbg.py 2>&1 | tail -14
