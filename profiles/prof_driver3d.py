import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleqp_b200 import Fact, problems
p = problems.poisson_control(int(sys.argv[1]) if len(sys.argv) > 1 else 32, 3)
f = Fact(device=0)
f.set_matrix(p.N, *p.kkt_lower())
print(f.stats()["ms_numeric"])
f.release()
