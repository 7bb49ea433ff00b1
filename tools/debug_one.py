import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.debug_parity import run
run(8, 512, 256, 1, 64, sys.argv[1] if len(sys.argv) > 1 else "fp32_simt")
