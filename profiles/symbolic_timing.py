import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleqp_b200 import Symbolic, problems
for cfg in (1,2):
    p = problems.config(cfg)
    cp, ri, v = p.kkt_lower()
    t=time.time(); s = Symbolic(p.N, cp, ri, v); print(p.name, 'analyze', round((time.time()-t)*1e3), 'ms', s.stats()['ms_symbolic'])
