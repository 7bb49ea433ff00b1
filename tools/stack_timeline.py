"""In-kernel timeline of the persistent whole-stack kernel (clock64 per phase; nf_debug_tc_timeline)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nf_b200
from nf_b200 import _lib
from oracle import synth
lib = _lib.lib
lib.nf_debug_tc_timeline.argtypes = [ctypes.c_void_p]
name = sys.argv[1] if len(sys.argv) > 1 else "C2a"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
cfg = dict(synth.CONFIGS[name]); eps = cfg.pop("ln_eps"); H1 = cfg.pop("H1")
dev = torch.device("cuda:0")
sd = synth.make_state_dict(cfg["D"], cfg["C"], cfg["R"], cfg["n"], seed=0, H1=H1)
m = nf_b200.RealNVP(cfg["D"], cfg["C"], cfg["R"], cfg["n"], None, first_hidden=H1, ln_eps=eps)
m.load_state_dict({k: torch.tensor(v) for k, v in sd.items()}); m = m.to(dev)
x, y = synth.make_inputs(B, cfg["D"], cfg["R"])
x = torch.tensor(x, device=dev); y = torch.tensor(y, device=dev)
lib.nf_graphs_enable(0)
with torch.no_grad():
    for _ in range(3): m.reverse(x, y)
    torch.cuda.synchronize()
    n = cfg["n"]; nph = 2 * n
    ctas = ((B + 127) // 128) * (cfg["C"] // 128) * 2
    buf = torch.zeros(ctas * nph * 12, dtype=torch.int64, device=dev)
    lib.nf_debug_tc_timeline(buf.data_ptr())
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); m.reverse(x, y); e1.record(); torch.cuda.synchronize()
    lib.nf_debug_tc_timeline(None)
t = buf.view(ctas, nph, 12).cpu().double()
NAMES = ["wait first tile", "main loop (split)", "mma drain", "tmem->regs + bias/lrelu", "stats exchange", "normalise + stores issued (+W3 partials)",
         "mask/rstd + tail", "bulk stores complete", "fences", "cluster barrier"]
print(f"{name} B={B}: reverse {e0.elapsed_time(e1) * 1e3:.1f} us (event), {ctas} CTAs, {nph} phases; mean cycles per phase segment (CTA-mean [min..max])")
for ph in (0, 1):
    sel = t[:, ph::2, :]
    sel = sel[:, 1:, :] if sel.shape[1] > 1 else sel     # skip the first block (cold)
    print(f"  phase {ph} ({'L1' if ph == 0 else 'L2 + tail'}):")
    for i, nm in enumerate(NAMES):
        d = sel[:, :, i + 1] - sel[:, :, i]
        print(f"    {nm:44s} {d.mean():8.0f}  [{d.min():7.0f} .. {d.max():7.0f}]")
    if ph == 1:
        lead = t[0:1, 3::2, :]          # CTA (0,0,0) is a cluster leader
        print(f"    {'leader: stores issued -> partials arrived':44s} {(lead[:, :, 11] - lead[:, :, 6]).mean():8.0f}")
        print(f"    {'leader: partials arrived -> tail done':44s} {(lead[:, :, 7] - lead[:, :, 11]).mean():8.0f}")
    tot = sel[:, :, 10] - sel[:, :, 0]
    print(f"    {'phase total':44s} {tot.mean():8.0f}  = {tot.mean() / 1.965e3:.2f} us")
whole = t[:, -1, 10] - t[:, 0, 0]
print(f"  whole stack in CTA: {whole.mean():.0f} cycles = {whole.mean() / 1.965e3:.1f} us")
