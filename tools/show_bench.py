"""Print the headline fields of a bench.py JSON line: python tools/show_bench.py FILE"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4g %s | e2e %.4g | roofline frac %.3f (%s) | cpu baseline %.3g on %d cores | clocks %s" % (
    d["value"], d["unit"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel"], (d.get("cpu_baseline") or {"value": float("nan")})["value"],
    (d.get("cpu_baseline") or {"cores": 0})["cores"], d["clocks"]))
for k, v in d.get("extra", {}).items():
    if k == "sweep":
        for kk, vv in v.items():
            print("  sweep %-22s %d envs per GPU: env path %.3e (frac %.3f per GPU), DQN in the loop %.3e" % (
                kk, vv["envs_per_gpu"], vv["value_env_only"], vv["roofline_frac_per_gpu_env_only"], vv["value_dqn_loop"]))
        continue
    print("  %-28s %s" % (k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "roofline_frac", "value_cuda_graph", "add_transitions_per_s", "error")}))
