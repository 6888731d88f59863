"""8 heated chains on rbcl738 through the C++ host driver on the visible GPU(s): in turn against side by side
(bench.py's mc3 block on its own)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from bench import mc3_block
n = torch.cuda.device_count()
devs = list(range(n)) if len(sys.argv) < 2 else [int(x) for x in sys.argv[1].split(",")]
print(json.dumps(mc3_block(devs)[0], indent=1))
