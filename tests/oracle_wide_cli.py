"""Runs the WIDE CPU oracle (oracle/liboracle_w12.so, rows of 12 words) in its own process and prints JSON:
    SGS_ORACLE_OW=12 python tests/oracle_wide_cli.py klocal <file> | periodic <file> <cmax> <c>
TEST INFRASTRUCTURE ONLY (tests/test_gpu_klocal_wide.py); oracle_py fixes its structure sizes at import, hence the
separate process."""
import json
import os
import sys

os.environ['SGS_ORACLE_OW'] = '12'
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_py as O  # noqa: E402

if sys.argv[1] == 'klocal':
    r = O.klocal(path=sys.argv[2], threads=os.cpu_count() or 1)
else:
    r = O.periodic(sys.argv[2], cmax=int(sys.argv[3]), c=int(sys.argv[4]))
print(json.dumps(r))
