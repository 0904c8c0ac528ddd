import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: round(d[k],3) for k in ("value","ms_per_step","factor_ms","solve_ms")}, {k: (round(v,3) if isinstance(v,float) else v) for k,v in d["e2e"].items()}, {k: round(v,2) for k,v in d["roofline"]["step_share_ms"].items()}, "frac", round(d["roofline"]["frac"],4))
    except Exception as e:
        print(f, "ERR", e)
