"""fuzz_sed: see tests/_fuzz_sed.py (python scratch/fuzz_sed.py [seed] [cases])."""
import sys
sys.path[:0] = ['.', 'tests']
from _fuzz_sed import run
failures = run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 60)
print('fuzz done, failures:', len(failures))
