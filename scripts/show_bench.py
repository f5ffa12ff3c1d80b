"""Condensed view of one bench.py JSON line (last line of the file given)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
def show(name, r):
    if "error" in r:
        print(name, "ERROR", r["error"]); return
    rl = r["roofline"]
    print(f"{name:14s} ms {r['ms_per_step']:.5f} (min {r.get('ms_per_step_min', 0):.5f}) value {r['value']:.1f} {r['unit']} "
          f"frac {rl['frac']:.3f} blocks {r.get('blocks')} e2e {r['e2e']['value']:.1f} (graph {r['e2e'].get('cuda_graph')}, "
          f"match {r['e2e'].get('output_matches_resident_run')}) clk {r['clocks']['sm_mhz']} {r['clocks']['reasons']} "
          f"n{r['clocks']['samples']} {r.get('run', {}).get('kernel')} R{r.get('run', {}).get('rotate_caches')} "
          f"parity {r.get('parity_check')}")
show("HEAD " + d["config"]["workload"][:8], d)
for k, v in d.get("workloads", {}).items():
    show(k, v)
print("scaling", d["scaling"], "n_gpus", d["n_gpus"], "cpu", d.get("cpu_baseline"), "wall", d.get("wall_s"))
