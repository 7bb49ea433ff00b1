"""Print the per-config table of a bench.py JSON line (python tools/bench_summary.py file.json)."""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("N=%d headline %s: value %.3e %s (%.4f ms) | e2e %.3e (%.4f ms) | sampling %.3e | launches/step %d" % (
    d["n_gpus"], d["config"]["workload"].split(":")[0], d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"],
    d["e2e"]["ms_per_step"], d["sampling"]["value"], d["gpu_launches_per_step"]))
print("  clocks", d["clocks"], "| roofline %.1f TF/s frac %.4f (%s) whole-step %.1f" % (
    d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["kernel"][:24], d["roofline"]["whole_step"]["achieved"]))
if d.get("cpu_baseline"):
    print("  cpu %.3e %s (%d cores)" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["unit"], d["cpu_baseline"]["cores"]))
for c, e in d["per_config"].items():
    row = "%-4s %-6s B/gpu %-6d" % (c, e["scaling"], e["config"]["batch_per_gpu"])
    if "train" in e:
        row += " train %.3e/s %8.3f ms" % (e["train"]["value"], e["train"]["ms_per_step"])
    row += " | sample %.3e/s %8.3f ms" % (e["sampling"]["value"], e["sampling"]["ms_per_step"])
    if "train" in e["e2e"]:
        row += " | e2e train %8.3f ms" % e["e2e"]["train"]["ms_per_step"]
    row += " e2e sample %8.3f ms" % e["e2e"]["sampling"]["ms_per_step"]
    r = e["roofline"]
    row += " | %s %.1f TF/s frac %.4f whole %.1f" % (r["kernel"].split(" ")[0][:10], r["achieved"], r["frac"], r["whole_step"]["achieved"])
    if e.get("cpu_baseline"):
        row += " | cpu %.3e" % e["cpu_baseline"]["value"]
    print(row)
    print("      breakdown", {k: round(v["ms_per_step"], 3) for k, v in r["breakdown"].items()})
    if "tf32" in e:
        t = e["tf32"]
        print("      tf32:", " ".join("%s %.3e/s %.3f ms" % (k, t[k]["value"], t[k]["ms_per_step"]) for k in ("train", "sampling") if k in t),
              "| err", {k: "%.1e" % v["err"] for k, v in t["error_vs_float64"].items() if isinstance(v, dict)},
              "| whole %.1f TF/s" % t["roofline"]["whole_step"]["achieved"])
