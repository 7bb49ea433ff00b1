"""Regenerates the measured tables of DESIGN.md (between the <!-- bench:begin --> / <!-- bench:end --> markers and the
c2t / affine markers) from a bench.py JSON line and a tools/affine_bw.py log, and keeps the line under profiles/.

    python tools/bench_table.py gpurun_out/r2h_bench.json gpurun_out/r2h_affine_bw.log
"""
import json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
aff = open(sys.argv[2]).read().strip().splitlines() if len(sys.argv) > 2 else []
with open(os.path.join(ROOT, "profiles", "r2_bench_line.json"), "w") as f:
    json.dump(d, f, indent=1)

def M(v): return f"{v / 1e6:.2f} M" if v >= 1e6 else f"{v / 1e3:.1f} k"
rows = ["| config (batch / GPU) | train step | train samples/s | sampling | sampled actions/s | e2e train / sampling | step TFLOP/s (of 273) | GEMM class TFLOP/s | CPU baseline (16 threads) | × CPU (e2e) |",
        "|---|---|---|---|---|---|---|---|---|---|"]
for name, v in d["per_config"].items():
    tr, sa, e = v.get("train"), v["sampling"], v["e2e"]
    cpu = v.get("cpu_baseline", {})
    roof = v["roofline"]
    e2e_main = e.get("train", e.get("sampling"))["value"]
    ratio = e2e_main / cpu["value"] if cpu.get("value") else None
    rows.append(f"| {name} ({v['config']['batch_per_gpu']}) | {tr['ms_per_step']:.3f} ms | {M(tr['value'])} |" if tr else f"| {name} ({v['config']['batch_per_gpu']}) | — | — |")
    rows[-1] += (f" {sa['ms_per_step']:.3f} ms | {M(sa['value'])} | {M(e['train']['value']) if 'train' in e else '—'} / {M(e['sampling']['value'])} |"
                 f" {roof['whole_step']['achieved']:.1f} | {roof['achieved']:.1f} | {M(cpu['value']) + ' ' + cpu['unit'] if cpu else '—'} | {ratio:.0f}× |" if ratio else " — |")
    if "tf32" in v:
        t = v["tf32"]
        rows.append(f"| {name}, `precision=\"tf32\"` (z err {t['error_vs_float64']['z']['err']:.1e}) | "
                    + (f"{t['train']['ms_per_step']:.3f} ms | {M(t['train']['value'])} |" if "train" in t else "— | — |")
                    + f" {t['sampling']['ms_per_step']:.3f} ms | {M(t['sampling']['value'])} | — | {t['roofline']['whole_step']['achieved']:.1f} | {t['roofline']['achieved']:.1f} | — | — |")
head = (f"Headline (`value`): C2a, {d['steps']} timed steps: **{d['ms_per_step']:.3f} ms per step = {M(d['value'])} samples/s** "
        f"(round 1: 0.740 ms / 5.53 M), sampling {M(d['sampling']['value'])} actions/s; end to end (pinned H2D + loss D2H every step) "
        f"**{M(d['e2e']['value'])} samples/s** ({d['e2e']['ms_per_step']:.3f} ms; with the library's fused `model.nll` loss "
        f"{M(d['e2e']['with_fused_loss']['value'])}); CPU arm {M(d['cpu_baseline']['value'])} samples/s on {d['cpu_baseline']['cores']} host threads; "
        f"{d['gpu_launches_per_step']} library kernels per step inside replayed graphs; clocks {d['clocks']['sm_mhz']:.0f} MHz, reasons {d['clocks']['reasons']}.  "
        f"Dominant class {d['roofline']['kernel']}: {d['roofline']['achieved']:.1f} TFLOP/s = {d['roofline']['frac']:.3f} of the measured bf16 peak "
        f"({d['roofline']['frac'] * 6:.2f} of the 3×TF32 ceiling), {d['roofline']['share_of_step'] * 100:.0f} % of the summed kernel time; "
        f"per-class ms/step: " + ", ".join(f"{k} {x['ms_per_step']:.3f}" for k, x in d["roofline"]["breakdown"].items()) + ".")
table = head + "\n\n" + "\n".join(rows) + "\n\n(one B200, L2 flushed between timed steps, `dy` included; C4 / C5 run their full BASELINE batch of 65 536 on the one GPU; " \
        "the e2e ratio divides by the CPU baseline of the same config, measured on a bounded sample — `cpu_baseline.sample` in the line.)"
c2t = d["per_config"].get("C2t")
c2t_txt = ""
if c2t:
    b = c2t["roofline"].get("breakdown") or {}
    c2t_txt = (f"train step {c2t['train']['ms_per_step']:.2f} ms = {M(c2t['train']['value'])} samples/s ({c2t['roofline']['whole_step']['achieved']:.0f} TFLOP/s "
               f"algorithmic over the whole step, GEMM class {c2t['roofline']['achieved']:.0f}), {c2t['gpu_launches_per_step']} kernels per step; autoregressive sampling "
               f"{c2t['sampling']['ms_per_step']:.2f} ms = {M(c2t['sampling']['value'])} actions/s; end to end {M(c2t['e2e']['train']['value'])} / {M(c2t['e2e']['sampling']['value'])}; "
               f"the checker on the 16 host threads (batch 256): {M(c2t['cpu_baseline']['value'])} samples/s and {M(c2t['cpu_baseline']['sampling']['value'])} actions/s"
               + ("; per-class ms/step: " + ", ".join(f"{k} {x['ms_per_step']:.2f}" for k, x in b.items()) if b else "") + ".")
aff_txt = "; ".join(re.sub(r"^affine ", "", l).replace(" of measured HBM peak", "") for l in aff) or "(not measured)"
p = os.path.join(ROOT, "DESIGN.md")
s = open(p).read()
def put(s, tag, txt):
    a, b = f"<!-- {tag}:begin -->", f"<!-- {tag}:end -->"
    if a in s:
        return s[:s.index(a) + len(a)] + "\n" + txt + "\n" + s[s.index(b):]
    return s.replace(f"@@{tag.upper()}@@", a + "\n" + txt + "\n" + b)
s = put(s, "bench_table", table); s = put(s, "c2t", c2t_txt); s = put(s, "affine", aff_txt)
open(p, "w").write(s)
print(table); print(c2t_txt); print(aff_txt)
