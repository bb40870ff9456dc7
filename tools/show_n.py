import json,sys
for l in open(sys.argv[1]):
    try: r=json.loads(l)
    except Exception: continue
    rs={r["config"]["workload"][:10]: r}; rs.update(r.get("workloads",{}))
    for k,v in rs.items():
        it=v.get("iteration") or {}
        e=v.get("e2e") or {}
        print(k, "value", round(v["value"],1), "ms", round(v["ms_per_step"],3), "| nccl-iter ms", round(it.get("ms_per_step",0),3), "ag_alone", round(it.get("allgather_alone_ms",0),3), "| fused", json.dumps({a:(round(b.get("ms_per_step"),3), b.get("all_ranks_hold_identical_result")) for a,b in (it.get("fused") or {}).items() if isinstance(b,dict) and "ms_per_step" in b}), "| e2e", round(e.get("value",0),1), round(e.get("ms_per_step",0),2))
